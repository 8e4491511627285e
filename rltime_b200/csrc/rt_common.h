// Shared host-side helpers for the rltime_b200 C-ABI library.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "rltime_b200.h"

namespace rt {

inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

inline std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> c{0};
  return c;
}

#define RT_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return rt::fail(RT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
                      cudaGetErrorString(e__));                                    \
  } while (0)

#define RT_LAUNCH_CHECK()                                                          \
  do {                                                                             \
    rt::launch_counter()++;                                                        \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess)                                                        \
      return rt::fail(RT_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, \
                      cudaGetErrorString(e__));                                    \
  } while (0)

#define RT_REQUIRE(cond, ...)                                                      \
  do {                                                                             \
    if (!(cond)) return rt::fail(RT_ERR_INVALID, __VA_ARGS__);                     \
  } while (0)

template <typename T>
inline cudaError_t dmalloc(T** p, size_t count) {
  return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

}  // namespace rt
