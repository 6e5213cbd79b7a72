// common.h -- error handling, dtype/context enums shared by the host runtime of libb200kv.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200kv {

// Synchronous failures become -1 + MXGetLastError() at the C boundary, like dmlc::Error raised by
// CHECK/LOG(FATAL) in the reference (include/mxnet/c_api_error.h:36-58).
struct Error : public std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

class ErrStream {
 public:
  ErrStream(const char* file, int line) { os_ << "[" << file << ":" << line << "] "; }
  template <typename T>
  ErrStream& operator<<(const T& v) {
    os_ << v;
    return *this;
  }
  [[noreturn]] void Throw() { throw Error(os_.str()); }
  std::ostringstream os_;
};

struct ErrThrower {
  [[noreturn]] void operator&(ErrStream& s) { s.Throw(); }
};

#define KV_FILE_ (__builtin_strrchr(__FILE__, '/') ? __builtin_strrchr(__FILE__, '/') + 1 : __FILE__)
#define KV_FATAL ::b200kv::ErrThrower() & ::b200kv::ErrStream(KV_FILE_, __LINE__)
#define KV_CHECK(cond) \
  if (!(cond)) KV_FATAL << "Check failed: " #cond << " "
#define KV_CHECK_EQ(a, b) \
  if (!((a) == (b))) KV_FATAL << "Check failed: " #a " == " #b << " (" << (a) << " vs. " << (b) << ") "
#define KV_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      KV_FATAL << "CUDA: " << cudaGetErrorName(e_) << ": " << cudaGetErrorString(e_) << " in " #call; \
  } while (0)

// include/mxnet/base.h:104-109
enum DevType { kCPU = 1, kGPU = 2, kCPUPinned = 3, kCPUShared = 5 };
// 3rdparty/mshadow/mshadow/base.h:307-314 (+ bf16 extension)
enum DType { kFloat32 = 0, kFloat64 = 1, kFloat16 = 2, kUint8 = 3, kInt32 = 4, kInt8 = 5, kInt64 = 6,
             kBool = 7, kBfloat16 = 12 };
// include/mxnet/ndarray.h:61-66
enum StorageType { kUndefinedStorage = -1, kDefaultStorage = 0, kRowSparseStorage = 1, kCSRStorage = 2 };

inline size_t DTypeSize(int dtype) {
  switch (dtype) {
    case kFloat32: case kInt32: return 4;
    case kFloat64: case kInt64: return 8;
    case kFloat16: case kBfloat16: return 2;
    case kUint8: case kInt8: case kBool: return 1;
  }
  KV_FATAL << "unknown dtype " << dtype;
}

inline const char* DTypeName(int dtype) {
  switch (dtype) {
    case kFloat32: return "float32"; case kFloat64: return "float64"; case kFloat16: return "float16";
    case kUint8: return "uint8"; case kInt32: return "int32"; case kInt8: return "int8";
    case kInt64: return "int64"; case kBool: return "bool"; case kBfloat16: return "bfloat16";
  }
  return "unknown";
}

struct Context {
  int dev_type = kCPU;
  int dev_id = 0;
  bool is_gpu() const { return dev_type == kGPU; }
  bool operator==(const Context& o) const { return dev_type == o.dev_type && dev_id == o.dev_id; }
  bool operator!=(const Context& o) const { return !(*this == o); }
  static Context GPU(int id) { return Context{kGPU, id}; }
  static Context CPU() { return Context{kCPU, 0}; }
  static Context Pinned() { return Context{kCPUPinned, 0}; }
  std::string str() const {
    std::ostringstream os;
    os << (dev_type == kGPU ? "gpu" : dev_type == kCPUPinned ? "cpu_pinned" : "cpu") << "(" << dev_id << ")";
    return os.str();
  }
};

constexpr int kMaxDevices = 16;

}  // namespace b200kv
