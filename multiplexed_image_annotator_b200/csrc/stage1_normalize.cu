// Stage 1: per-channel image normalisation.  Replaces ImageProcessor._normalize
// (reference cta/preprocess.py:214-239): sigma-20 Gaussian background subtraction (bg clamped to
// 125), optional Gaussian blur, upper-percentile clip, 2*x/max(25,max)-1.
//
// Bit-exact with scipy/numpy by construction:
//   * the separable Gaussians follow scipy.ndimage.correlate1d's symmetric path operation by
//     operation: float64 accumulate  acc = x[c]*w0 ; for k = r..1: acc += (x[c-k] + x[c+k]) * w[k]
//     with separate multiply and add, 'reflect' boundary, float32 store after each axis (axis 0
//     first).  The tap weights come from the host (numpy's exp), so they are scipy's to the bit.
//   * the percentile is numpy's 'linear' rule on the exact order statistics k_lo / k_hi, found by a
//     3-pass radix select over the float32 bit patterns (11 + 11 + 10 bits).
//
// FIR kernels: one 64 x 32 output tile per CTA; the tile plus its +-r halo is staged in shared
// memory as float64 (coalesced global reads, one conversion per element); every thread produces 8
// consecutive outputs ALONG the filter axis and keeps two sliding 15-element register windows, so a
// group of 8 taps x 8 outputs costs 30 shared loads for 192 FP64 operations (FP64-pipe bound, not
// LSU bound).  The plane of one channel (67 MB at 4096^2) stays L2-resident between the passes.
#include "common.cuh"

namespace ribca {

struct Taps {
  double w[RIBCA_MAX_TAPS + 1];   // w[k] = weight at distance k
  int r;
};

enum { EPI_STORE = 0, EPI_BG = 1 };

struct SelectState {          // lives in the workspace, one per call (channels are serialised)
  unsigned int hist[2][2048];
  unsigned int prefix[2];     // bits fixed so far for the two ranks
  long long rank[2];          // remaining rank inside the current prefix bucket
  int max_bits;               // float bits of max(x) (x >= 0 so int order == float order)
};

constexpr int kFirThreads = 256;
constexpr int kOutF = 64;     // outputs per tile along the filter axis (8 thread groups x 8)
constexpr int kOutL = 32;     // outputs per tile along the other axis (one per lane)
constexpr int kPitch = 33;    // shared row pitch in doubles (odd: conflict-free transposed stores)

__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy 'reflect' (d c b a | a b c d | d c b a), valid for any offset
  const int p = 2 * n;
  int m = i % p;
  if (m < 0) m += p;
  return m < n ? m : p - 1 - m;
}

template <typename T> __device__ __forceinline__ float load_as_float(const T* p, long long i);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p, long long i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float load_as_float<uint16_t>(const uint16_t* p, long long i) { return (float)__ldg(p + i); }
template <> __device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p, long long i) { return (float)__ldg(p + i); }
template <> __device__ __forceinline__ float load_as_float<int32_t>(const int32_t* p, long long i) { return (float)__ldg(p + i); }

// HORIZ = false: filter along rows (axis 0), lanes along columns.
// HORIZ = true : filter along columns (axis 1), lanes along rows.
template <typename Tin, typename Traw, bool HORIZ, int EPI, bool STATS>
__global__ void __launch_bounds__(kFirThreads)
fir_kernel(const Tin* __restrict__ in, const Traw* __restrict__ raw, float* __restrict__ out, int H, int W,
           const __grid_constant__ Taps taps, SelectState* st) {
  extern __shared__ double tile[];               // (kOutF + 2r) x kPitch
  __shared__ double w_s[RIBCA_MAX_TAPS + 1];
  __shared__ int red_max[kFirThreads / 32];
  const int r = taps.r;
  const int F = kOutF + 2 * r;
  const int tid = threadIdx.x;
  for (int k = tid; k <= r; k += kFirThreads) w_s[k] = taps.w[k];

  const int nF = HORIZ ? W : H;                   // extent along the filter axis
  const int nL = HORIZ ? H : W;
  const int f0 = blockIdx.x * kOutF;              // tile origin along the filter axis
  const int l0 = blockIdx.y * kOutL;

  // ---- stage tile + halo (float64) --------------------------------------------------------------
  if (!HORIZ) {
    const int l = tid & 31;
    const int gl = l0 + l;
    for (int f = tid >> 5; f < F; f += kFirThreads / 32) {
      const int gf = reflect_index(f0 - r + f, nF);
      double v = 0.0;
      if (gl < nL) v = (double)load_as_float<Tin>(in, (long long)gf * W + gl);
      tile[f * kPitch + l] = v;
    }
  } else {
    for (int idx = tid; idx < F * kOutL; idx += kFirThreads) {
      const int l = idx / F, f = idx - l * F;
      const int gl = l0 + l;
      const int gf = reflect_index(f0 - r + f, nF);
      double v = 0.0;
      if (gl < nL) v = (double)load_as_float<Tin>(in, (long long)gl * W + gf);
      tile[f * kPitch + l] = v;
    }
  }
  __syncthreads();

  // ---- 8 consecutive outputs along the filter axis per thread -------------------------------------
  const int lane = tid & 31;
  const int grp = tid >> 5;                       // 0..7
  const double* col = tile + lane;                // element f at col[f * kPitch]
  const int c = r + grp * 8;                      // tile index of this thread's first output centre
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = __dmul_rn(col[(c + i) * kPitch], w_s[0]);

  int k = r;
  // leading taps (r % 8 of them), straightforward
  for (int rem = r & 7; rem > 0; --rem, --k) {
    const double wk = w_s[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double pair = __dadd_rn(col[(c + i - k) * kPitch], col[(c + i + k) * kPitch]);
      acc[i] = __dadd_rn(acc[i], __dmul_rn(pair, wk));
    }
  }
  // groups of 8 taps with sliding register windows:
  //   left  value for output i at tap k-u : x[c - k + (i + u)]      -> L[i + u],      L[m] = x[c - k + m]
  //   right value for output i at tap k-u : x[c + k - 7 + (i - u + 7)] -> R[i - u + 7], R[m] = x[c + k - 7 + m]
  for (; k >= 8; k -= 8) {
    double L[15], R[15];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      L[m] = col[(c - k + m) * kPitch];
      R[m + 7] = col[(c + k + m) * kPitch];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u > 0) {
        L[u + 7] = col[(c - k + u + 7) * kPitch];
        R[7 - u] = col[(c + k - u) * kPitch];
      }
      const double wk = w_s[k - u];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double pair = __dadd_rn(L[i + u], R[i - u + 7]);
        acc[i] = __dadd_rn(acc[i], __dmul_rn(pair, wk));
      }
    }
  }

  // ---- epilogue -----------------------------------------------------------------------------------
  const int gl = l0 + lane;
  float local_max = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gf = f0 + grp * 8 + i;
    if (gl < nL && gf < nF) {
      const long long o = HORIZ ? ((long long)gl * W + gf) : ((long long)gf * W + gl);
      float v = __double2float_rn(acc[i]);
      if (EPI == EPI_BG) {
        const float bg = v > 125.0f ? 125.0f : v;
        v = fmaxf(__fsub_rn(load_as_float<Traw>(raw, o), bg), 0.0f);
      }
      out[o] = v;
      local_max = fmaxf(local_max, v);
    }
  }
  if (STATS) {
    local_max = warp_max(local_max);
    if (lane == 0) red_max[grp] = __float_as_int(local_max);
    __syncthreads();
    if (tid == 0) {
      int m = red_max[0];
#pragma unroll
      for (int g = 1; g < kFirThreads / 32; ++g) m = max(m, red_max[g]);
      if (m > 0) atomicMax(&st->max_bits, m);
    }
  }
}

// ---- radix select -----------------------------------------------------------------------------
__global__ void select_init_kernel(SelectState* st, long long k_lo, long long k_hi) {
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&st->hist[0][0])[i] = 0u;
  if (threadIdx.x == 0) {
    st->prefix[0] = st->prefix[1] = 0u;
    st->rank[0] = k_lo;
    st->rank[1] = k_hi;
    st->max_bits = 0;
  }
}

__device__ __forceinline__ unsigned int float_key(float v) {
  unsigned int b = __float_as_uint(v);
  return b == 0x80000000u ? 0u : b;       // -0.0 sorts with +0.0 (values are >= 0 here)
}

template <int PASS>
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ x, long long n, SelectState* st) {
  constexpr int kShift = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
  constexpr int kBits = PASS == 2 ? 10 : 11;
  constexpr int kHiShift = kShift + kBits;
  __shared__ unsigned int h[2][2048];
  for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&h[0][0])[i] = 0u;
  __syncthreads();
  const unsigned int p0 = st->prefix[0], p1 = st->prefix[1];
  const bool same = (PASS == 0) || (p0 == p1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_round = (n + 31) / 32 * 32;   // keep warps converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    int bin0 = -1, bin1 = -1;
    if (i < n) {
      const unsigned int key = float_key(__ldg(x + i));
      const unsigned int hi = PASS == 0 ? 0u : (key >> kHiShift);
      const int bin = (int)((key >> kShift) & ((1u << kBits) - 1u));
      if (PASS == 0 || hi == p0) bin0 = bin;
      if (!same && hi == p1) bin1 = bin;
    }
    // warp-aggregated shared atomics: one add per distinct bin per warp
    unsigned int peers = __match_any_sync(0xffffffffu, bin0);
    if (bin0 >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&h[0][bin0], __popc(peers));
    if (!same) {
      peers = __match_any_sync(0xffffffffu, bin1);
      if (bin1 >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&h[1][bin1], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
    if (h[0][i]) atomicAdd(&st->hist[0][i], h[0][i]);
    if (!same && h[1][i]) atomicAdd(&st->hist[1][i], h[1][i]);
  }
}

template <int PASS>
__global__ void __launch_bounds__(1024) select_pick_kernel(SelectState* st) {
  constexpr int kBits = PASS == 2 ? 10 : 11;
  __shared__ int total;
  __shared__ unsigned int new_prefix[2];
  __shared__ long long new_rank[2];
  const bool same = (PASS == 0) || (st->prefix[0] == st->prefix[1]);
  const unsigned int old_prefix[2] = {st->prefix[0], st->prefix[1]};
  const long long old_rank[2] = {st->rank[0], st->rank[1]};
  __syncthreads();
  for (int q = 0; q < 2; ++q) {
    const unsigned int* hist = st->hist[(same ? 0 : q)];
    const int b0 = 2 * threadIdx.x, b1 = b0 + 1;
    const int c0 = (int)hist[b0], c1 = (int)hist[b1];     // counts fit int32 per bin only if n < 2^31
    const int ex = block_exclusive_scan(c0 + c1, &total);
    const long long rk = old_rank[q];
    if (rk >= ex && rk < (long long)ex + c0) {
      new_prefix[q] = (old_prefix[q] << kBits) | (unsigned int)b0;
      new_rank[q] = rk - ex;
    } else if (rk >= (long long)ex + c0 && rk < (long long)ex + c0 + c1) {
      new_prefix[q] = (old_prefix[q] << kBits) | (unsigned int)b1;
      new_rank[q] = rk - ex - c0;
    }
    __syncthreads();
  }
  // clear the histograms for the next pass / next channel
  st->hist[0][2 * threadIdx.x] = 0u; st->hist[0][2 * threadIdx.x + 1] = 0u;
  st->hist[1][2 * threadIdx.x] = 0u; st->hist[1][2 * threadIdx.x + 1] = 0u;
  if (threadIdx.x < 2) {
    st->prefix[threadIdx.x] = new_prefix[threadIdx.x];
    st->rank[threadIdx.x] = new_rank[threadIdx.x];
  }
}

// ---- final affine map -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_apply_kernel(float* __restrict__ x, long long n, const SelectState* __restrict__ st, float gamma,
                       float* chan_stats) {
  // scalar plan, recomputed by every thread (numpy float32 arithmetic, preprocess.py:228-238)
  const float mx = __int_as_float(st->max_bits);
  const bool none_positive = !(mx > 0.0f);
  const float a = __uint_as_float(st->prefix[0]);
  const float b = __uint_as_float(st->prefix[1]);
  // numpy _lerp: a + (b - a) * t, and b - (b - a) * (1 - t) where t >= 0.5
  const float diff = __fsub_rn(b, a);
  float thresh = __fadd_rn(a, __fmul_rn(diff, gamma));
  if (gamma >= 0.5f) thresh = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
  const bool clip = thresh > 20.0f;
  const float mx_after = clip ? fminf(mx, thresh) : mx;
  const float denom = mx_after > 25.0f ? mx_after : 25.0f;
  if (chan_stats && blockIdx.x == 0 && threadIdx.x == 0) {
    chan_stats[0] = thresh;
    chan_stats[1] = mx_after;
    chan_stats[2] = none_positive ? 1.0f : 0.0f;
    chan_stats[3] = 0.0f;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = x[i];
    if (none_positive) {
      v = -1.0f;
    } else {
      if (clip) v = fminf(fmaxf(v, 0.0f), thresh);
      v = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(v, denom)), 1.0f);
    }
    x[i] = v;
  }
}

template <typename Tin, typename Traw, bool HORIZ, int EPI, bool STATS>
static int launch_fir(const Tin* in, const Traw* raw, float* out, int H, int W, const Taps& taps,
                      SelectState* st, cudaStream_t stream) {
  const int nF = HORIZ ? W : H, nL = HORIZ ? H : W;
  dim3 grid((nF + kOutF - 1) / kOutF, (nL + kOutL - 1) / kOutL);
  size_t smem = (size_t)(kOutF + 2 * taps.r) * kPitch * sizeof(double);
  auto kern = fir_kernel<Tin, Traw, HORIZ, EPI, STATS>;
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (int)((int)((kOutF + 2 * RIBCA_MAX_TAPS) * kPitch * sizeof(double))), "cudaFuncSetAttribute(fir_kernel)"));
  kern<<<grid, kFirThreads, smem, stream>>>(in, raw, out, H, W, taps, st);
  RIBCA_LAUNCH_CHECK("fir_kernel");
  return RIBCA_OK;
}

static int make_taps(Taps& t, const double* h_w, int r, const char* what) {
  RIBCA_REQUIRE(r >= 0 && r <= RIBCA_MAX_TAPS, "%s radius %d outside [0, %d]", what, r, RIBCA_MAX_TAPS);
  RIBCA_REQUIRE(h_w != nullptr, "%s weights are null", what);
  memset(&t, 0, sizeof(t));
  t.r = r;
  for (int k = 0; k <= r; ++k) t.w[k] = h_w[k];
  return RIBCA_OK;
}

template <typename Traw>
static int normalize_typed(const Traw* img, int C, int H, int W, const Taps& bg, const Taps* blur,
                           long long k_lo, long long k_hi, float gamma, float* out, float* chan_stats,
                           float* tmp, SelectState* st, cudaStream_t stream) {
  const long long hw = (long long)H * W;
  const int blocks = (int)std::min<long long>((hw + 255) / 256, (long long)num_sms() * 8);
  for (int c = 0; c < C; ++c) {
    const Traw* src = img + c * hw;
    float* dst = out + c * hw;
    select_init_kernel<<<1, 256, 0, stream>>>(st, k_lo, k_hi);
    RIBCA_LAUNCH_CHECK("select_init_kernel");
    RIBCA_TRY((launch_fir<Traw, Traw, false, EPI_STORE, false>(src, src, tmp, H, W, bg, st, stream)));
    if (blur) {
      RIBCA_TRY((launch_fir<float, Traw, true, EPI_BG, false>(tmp, src, dst, H, W, bg, st, stream)));
      RIBCA_TRY((launch_fir<float, Traw, false, EPI_STORE, false>(dst, src, tmp, H, W, *blur, st, stream)));
      RIBCA_TRY((launch_fir<float, Traw, true, EPI_STORE, true>(tmp, src, dst, H, W, *blur, st, stream)));
    } else {
      RIBCA_TRY((launch_fir<float, Traw, true, EPI_BG, true>(tmp, src, dst, H, W, bg, st, stream)));
    }
    select_hist_kernel<0><<<blocks, 256, 0, stream>>>(dst, hw, st);
    RIBCA_LAUNCH_CHECK("select_hist_kernel<0>");
    select_pick_kernel<0><<<1, 1024, 0, stream>>>(st);
    RIBCA_LAUNCH_CHECK("select_pick_kernel<0>");
    select_hist_kernel<1><<<blocks, 256, 0, stream>>>(dst, hw, st);
    RIBCA_LAUNCH_CHECK("select_hist_kernel<1>");
    select_pick_kernel<1><<<1, 1024, 0, stream>>>(st);
    RIBCA_LAUNCH_CHECK("select_pick_kernel<1>");
    select_hist_kernel<2><<<blocks, 256, 0, stream>>>(dst, hw, st);
    RIBCA_LAUNCH_CHECK("select_hist_kernel<2>");
    select_pick_kernel<2><<<1, 1024, 0, stream>>>(st);
    RIBCA_LAUNCH_CHECK("select_pick_kernel<2>");
    normalize_apply_kernel<<<blocks, 256, 0, stream>>>(dst, hw, st, gamma, chan_stats ? chan_stats + 4 * c : nullptr);
    RIBCA_LAUNCH_CHECK("normalize_apply_kernel");
  }
  return RIBCA_OK;
}

}  // namespace ribca

using namespace ribca;

extern "C" {

size_t ribca_normalize_workspace_bytes(int C, int H, int W) {
  (void)C;
  return align_up((size_t)H * W * sizeof(float), 256) + align_up(sizeof(SelectState), 256);
}

int ribca_normalize(const void* img, int dtype, int C, int H, int W, const double* h_w_bg, int r_bg,
                    const double* h_w_blur, int r_blur, long long k_lo, long long k_hi, float gamma,
                    float* out, float* chan_stats, void* workspace, size_t workspace_bytes,
                    ribca_stream_t stream) {
  RIBCA_REQUIRE(img && out && workspace, "ribca_normalize: null pointer");
  RIBCA_REQUIRE(C > 0 && H > 0 && W > 0, "ribca_normalize: bad shape C=%d H=%d W=%d", C, H, W);
  const long long hw = (long long)H * W;
  RIBCA_REQUIRE(hw < (1ll << 31), "ribca_normalize: plane of %lld pixels exceeds 2^31", hw);
  RIBCA_REQUIRE(k_lo >= 0 && k_lo < hw && k_hi >= 0 && k_hi < hw, "ribca_normalize: order statistics outside the plane");
  if (workspace_bytes < ribca_normalize_workspace_bytes(C, H, W)) {
    set_error("ribca_normalize: workspace %zu < %zu", workspace_bytes, ribca_normalize_workspace_bytes(C, H, W));
    return RIBCA_EWORKSPACE;
  }
  Taps bg, blur;
  RIBCA_TRY(make_taps(bg, h_w_bg, r_bg, "background"));
  const bool has_blur = r_blur >= 0;
  if (has_blur) RIBCA_TRY(make_taps(blur, h_w_blur, r_blur, "blur"));
  float* tmp = static_cast<float*>(workspace);
  SelectState* st = reinterpret_cast<SelectState*>(static_cast<char*>(workspace) + align_up((size_t)hw * sizeof(float), 256));
  cudaStream_t s = as_stream(stream);
  const bool prof = profiling();
  if (prof) {
    const double in_b = dtype == RIBCA_U8 ? 1.0 : (dtype == RIBCA_U16 ? 2.0 : 4.0);
    prof_begin_span(RIBCA_PROF_NORMALIZE, (double)C * (double)hw * (2.0 * in_b + 4.0), s);
  }
  int rc;
  switch (dtype) {
    case RIBCA_U8:
      rc = normalize_typed<uint8_t>(static_cast<const uint8_t*>(img), C, H, W, bg, has_blur ? &blur : nullptr, k_lo, k_hi, gamma, out, chan_stats, tmp, st, s);
      break;
    case RIBCA_U16:
      rc = normalize_typed<uint16_t>(static_cast<const uint16_t*>(img), C, H, W, bg, has_blur ? &blur : nullptr, k_lo, k_hi, gamma, out, chan_stats, tmp, st, s);
      break;
    case RIBCA_F32:
      rc = normalize_typed<float>(static_cast<const float*>(img), C, H, W, bg, has_blur ? &blur : nullptr, k_lo, k_hi, gamma, out, chan_stats, tmp, st, s);
      break;
    case RIBCA_I32:
      rc = normalize_typed<int32_t>(static_cast<const int32_t*>(img), C, H, W, bg, has_blur ? &blur : nullptr, k_lo, k_hi, gamma, out, chan_stats, tmp, st, s);
      break;
    default:
      set_error("ribca_normalize: unsupported dtype %d", dtype);
      rc = RIBCA_EUNSUPPORTED;
  }
  if (prof) prof_end_span(s);
  return rc;
}

}  // extern "C"
