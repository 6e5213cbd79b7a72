// op_params.h -- the string attribute dict of an imperative operator call, parsed the way the
// reference's dmlc::Parameter fields are (scalars: dmlc::stof; tuples: istream >> float).
#pragma once
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "common.h"
#include "scalar_parse.h"

namespace b200kv {

typedef std::vector<std::pair<std::string, std::string>> Params;

inline const std::string* Find(const Params& p, const std::string& k) {
  for (auto& kv : p) {
    if (kv.first == k) return &kv.second;
  }
  return nullptr;
}

// scalar op parameter: dmlc::Parameter float field (dmlc::stof)
inline float GetF(const Params& p, const std::string& k, float dflt) {
  const std::string* v = Find(p, k);
  return v ? DmlcStof(*v) : dflt;
}

inline int GetI(const Params& p, const std::string& k, int dflt) {
  const std::string* v = Find(p, k);
  return v ? std::atoi(v->c_str()) : dflt;
}

inline bool GetB(const Params& p, const std::string& k, bool dflt) {
  const std::string* v = Find(p, k);
  if (!v) return dflt;
  return *v == "True" || *v == "true" || *v == "1";
}

// tuple op parameter "(0.1, 0.2)" / "[0.1, 0.2]": mxnet::Tuple<float> via istream >> float
inline std::vector<float> GetTuple(const Params& p, const std::string& k) {
  const std::string* v = Find(p, k);
  KV_CHECK(v != nullptr) << "Required parameter " << k << " is missing";
  std::vector<float> out;
  const char* s = v->c_str();
  while (*s) {
    if (*s == '(' || *s == ')' || *s == '[' || *s == ']' || *s == ',' || *s == ' ' || *s == 'L') {
      ++s;
      continue;
    }
    char* end = nullptr;
    float f = std::strtof(s, &end);
    KV_CHECK(end != s) << "cannot parse tuple parameter " << k << "='" << *v << "'";
    out.push_back(f);
    s = end;
  }
  return out;
}


// integer tuple "(1, 2, 3)" (mxnet::Tuple<int>)
inline std::vector<int> GetIntTuple(const Params& p, const std::string& k) {
  std::vector<int> out;
  for (float f : GetTuple(p, k)) out.push_back(static_cast<int>(f));
  return out;
}

}  // namespace b200kv
