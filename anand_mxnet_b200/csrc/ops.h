// ops.h -- the imperative-operator surface (MXImperativeInvokeEx subset): the reference's optimizer
// operators (src/operator/optimizer_op.cc:195-710) plus the few elementwise helpers that updaters
// written in the host language rely on. All run on the arrays' GPU through the dense fused kernel.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "ndarray.h"

namespace b200kv {

struct OpInfo {
  std::string name;
};

// nullptr when the operator is not registered
const OpInfo* FindOp(const std::string& name);

// outputs: caller-provided (in place) or empty -> allocated and appended
void InvokeOp(const OpInfo* op, const std::vector<NDArray>& inputs, std::vector<NDArray>* outputs,
              const std::vector<std::pair<std::string, std::string>>& params);

// multi_ops.cc: the multi-tensor optimizer operators (multi_sum_sq, multi_lars, adamw, lamb);
// returns false when `name` is not one of them
bool MultiTensorOp(const std::string& name, const std::vector<NDArray>& inputs,
                   std::vector<NDArray>* outputs,
                   const std::vector<std::pair<std::string, std::string>>& params);

}  // namespace b200kv
