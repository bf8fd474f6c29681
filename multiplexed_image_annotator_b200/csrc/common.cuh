// Shared helpers for libribca_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
bool interleave_enabled();                 // RIBCA_INTERLEAVE / ribca_set_interleave, off while profiling
int side_stream(SideStream* out);          // this host thread's side stream on the current device (created on first use)

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

// ---- operand plane formats of the tensor-core contractions ---------------------------------------
// Every GEMM operand is two 16-bit planes [2][rows][K].
//   kFmtBf16  : {bf16 hi, bf16 lo}, x = hi + lo; product = lo.hi + hi.lo + hi.hi (three kind::f16 passes)
//   kFmtF16F8 : plane 0 = fp16(x); plane 1 = a PAIR of e4m3 per element that carries the two first-order
//               correction terms as one kind::f8f6f4 pass over a doubled K axis:
//                 A role: (e4m3((x - f16 x) * 2^8), e4m3(f16 x))
//                 W role: (e4m3(f16(w) * 2^t),      e4m3((w - f16 w) * 2^(t+8)))   and plane 0 = fp16(w * 2^(t+8))
//               so  A0.W0 + A1.W1 = 2^(t+8) * (a_hi w_hi + a_lo w_hi + a_hi w_lo) with the corrections rounded to
//               4 significant bits (relative error ~2^-15 of a product); the epilogue scales by 2^-(t+8).
//               Two tensor-core passes worth of time (the fp8 pass runs at twice the fp16 rate) instead of three.
constexpr int kFmtBf16 = 0;
constexpr int kFmtF16F8 = 1;
constexpr float kF16F8LoScale = 256.0f;       // 2^8

// d = {e4m3(hi_byte) << 8 | e4m3(lo_byte)}, round-to-nearest-even, saturating at +-448
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo_byte, float hi_byte) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi_byte), "f"(lo_byte));
  return (uint32_t)r;
}
// A-role split of two consecutive elements: p0 = f16x2(a, b), p1 = the two e4m3 pairs (element a in the low half).
// 9 instructions per pair: saturating f32x2 -> f16x2, f16x2 -> e4m3x2 (the hi8 copies), 2 unpacks, 2 FMA for the
// scaled remainders (a * 2^8 - h * 2^8, exact), f32x2 -> e4m3x2, 1 byte interleave.
__device__ __forceinline__ void split_f16f8_x2(float a, float b, uint32_t& p0, uint32_t& p1) {
  uint32_t h2;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(b), "f"(a));            // a in the low half
  p0 = h2;
  uint16_t hi8, lo8;
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(hi8) : "r"(h2));                    // byte 0 = e4m3(h.a), byte 1 = e4m3(h.b)
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
  const float la = fmaf(a, kF16F8LoScale, -hf.x * kF16F8LoScale);
  const float lb = fmaf(b, kF16F8LoScale, -hf.y * kF16F8LoScale);
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo8) : "f"(lb), "f"(la));        // byte 0 = e4m3(la), byte 1 = e4m3(lb)
  // bytes: [lo a, hi a, lo b, hi b]
  p1 = __byte_perm((uint32_t)lo8, (uint32_t)hi8, 0x5140);
}
// W-role split (weights, packed once): sc = 2^t
__device__ __forceinline__ void split_f16f8_w_x2(float a, float b, float sc, uint32_t& p0, uint32_t& p1) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const float s2 = sc * kF16F8LoScale;
  const __half2 m = __floats2half2_rn(hf.x * s2, hf.y * s2);        // exact power-of-two scaling of f16(w)
  p0 = *reinterpret_cast<const uint32_t*>(&m);
  p1 = cvt_e4m3x2(hf.x * sc, (a - hf.x) * s2) | (cvt_e4m3x2(hf.y * sc, (b - hf.y) * s2) << 16);
}
// two consecutive elements -> one 32-bit word per plane in the requested A-role format
__device__ __forceinline__ void split_pair(float a, float b, int fmt, uint32_t& p0, uint32_t& p1) {
  if (fmt == kFmtF16F8) split_f16f8_x2(a, b, p0, p1);
  else split_bf16x2(a, b, p0, p1);
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
