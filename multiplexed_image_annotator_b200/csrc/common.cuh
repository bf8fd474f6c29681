// Shared helpers for libribca_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "ribca_b200.h"

namespace ribca {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int ensure_dynamic_smem(const void* func, int bytes, const char* name);
bool profiling();
void prof_begin_span(int cls, double work, cudaStream_t st);
void prof_end_span(cudaStream_t st);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return RIBCA_ECUDA;
  }
  return RIBCA_OK;
}

// after a kernel launch: catches launch-configuration errors without synchronising
#define RIBCA_LAUNCH_CHECK(name)                                             \
  do {                                                                       \
    ::ribca::count_launch();                                                 \
    int _rc = ::ribca::check_cuda(cudaGetLastError(), name);                 \
    if (_rc != RIBCA_OK) return _rc;                                         \
  } while (0)

#define RIBCA_REQUIRE(cond, ...)                                             \
  do {                                                                       \
    if (!(cond)) {                                                           \
      ::ribca::set_error(__VA_ARGS__);                                       \
      return RIBCA_EINVAL;                                                   \
    }                                                                        \
  } while (0)

#define RIBCA_TRY(expr)                                                      \
  do {                                                                       \
    int _rc = (expr);                                                        \
    if (_rc != RIBCA_OK) return _rc;                                         \
  } while (0)

inline cudaStream_t as_stream(ribca_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits survive
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// (a, b) -> packed split bf16x2: hi = bf16x2(a, b), lo = bf16x2(a - hi.a, b - hi.b); element a in the low half.
// One packed conversion per pair instead of two scalar ones.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exclusive prefix sum over the block (blockDim.x multiple of 32, <= 1024); *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    warp_sums[lane] = wi - w;   // exclusive prefix of the warp totals
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  int res = warp_sums[warp] + incl - v;
  __syncthreads();
  return res;
}

}  // namespace ribca
