// Stage 4: the ViT classifiers and the MAE marker imputer, built on the two-plane tcgen05 GEMM.
// Replaces VisionTransformer / vit_* + softmax (reference cta/model.py:31-88, 397-406) and
// MaskedAutoencoderViT.forward / MarkerImputer.impute (cta/markerImputer.py:155-232, 294-329).
//
// Activations between kernels:  residual stream x fp32 [M][D] (M = cells * tokens);
// every GEMM A-operand is produced directly as two 16-bit planes (f16f8 or bf16 {hi, lo}, csrc/common.cuh)
// by the kernel before it (im2col, LayerNorm, attention, GELU epilogue), so no separate conversion pass
// touches HBM.  Attention runs on the tensor cores for the classifiers (stage4_attention.cu) and on the FP32
// pipe, with K/V staged in shared memory, for the imputer's 7-16-token sequences.
#include "common.cuh"

namespace ribca {

int gemm_launch(const void* A, long long a_plane, const void* W, long long w_plane, int M, int N, int K,
                const float* bias, const float* row_table, int table_period, int epilogue, float* out_f32,
                void* out_split, long long out_plane, int precision, int w_log2_scale, cudaStream_t stream,
                const ribca_ln_fold* ln = nullptr);
int gemm_ln_slots(int N, int precision);

int attention_tc_launch(const void* qkv_split, long long qkv_plane, int cells, int tokens, int heads, int hd,
                        void* out_split, long long out_plane, int out_fmt, cudaStream_t st);

int vit_forward_f32(const ribca_vit_desc* desc, const float* wf32, const float* wmat, const float* patches, int n_cells,
                    float* probs, float* logits, float* x, float* a, float* qkv, float* h, cudaStream_t st);

static inline int fmt_of(int precision) { return precision == RIBCA_F16F8 ? kFmtF16F8 : kFmtBf16; }

typedef __nv_bfloat16 bf16;

// ---- patches -> patch-embed A operand ---------------------------------------------------------
// A[(cell*101 + 1 + py*10 + px)][c*16 + ky*4 + kx] = patch[cell][c][py*4+ky][px*4+kx]; row cell*101 = 0
__global__ void __launch_bounds__(256)
im2col_split_kernel(const float* __restrict__ patches, int n_cells, int C, int fmt, bf16* __restrict__ a_hi,
                    bf16* __restrict__ a_lo) {
  const int Kpe = 16 * C;
  const long long total = (long long)n_cells * C * 40 * 10;       // one float4 (4 kx) per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int px = (int)(t % 10);
    const int y = (int)((t / 10) % 40);
    const int c = (int)((t / 400) % C);
    const long long cell = t / (400ll * C);
    const float4 v = __ldg(reinterpret_cast<const float4*>(patches) + t);
    const long long row = cell * 101 + 1 + (y >> 2) * 10 + px;
    const long long o = row * Kpe + c * 16 + (y & 3) * 4;
    uint32_t h[2], l[2];
    split_pair(v.x, v.y, fmt, h[0], l[0]);
    split_pair(v.z, v.w, fmt, h[1], l[1]);
    *reinterpret_cast<uint2*>(a_hi + o) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(a_lo + o) = make_uint2(l[0], l[1]);
  }
  // class-token rows are all-zero A rows (their value comes from the epilogue's row table)
  const long long ztotal = (long long)n_cells * Kpe;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ztotal; t += stride) {
    const long long cell = t / Kpe;
    const long long o = cell * 101 * Kpe + (t - cell * Kpe);
    a_hi[o] = __float2bfloat16_rn(0.0f);
    a_lo[o] = __float2bfloat16_rn(0.0f);
  }
}

// ---- LayerNorm -> split-bf16 ------------------------------------------------------------------
// one warp per row, D % 4 == 0, D <= 1024; fp32 two-pass statistics.  NI = float4 per lane (ceil(D / 128)): the row lives
// in 4 * NI registers, so that a 256-thread block stays under 48 registers per thread and two of them fit beside a
// resident GEMM CTA (the two-stream interleave, ribca_set_interleave).
template <int NI>
__global__ void __launch_bounds__(256, 5)
layernorm_split_kernel(const float* __restrict__ x, int M, int D, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, int fmt, bf16* __restrict__ o_hi, bf16* __restrict__ o_lo) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nv = D >> 2;                          // float4 per row
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * D);
    float4 v[NI];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) { v[i] = xr[idx]; sum += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + idx);
        uint32_t h[2], l[2];
        split_pair((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y, fmt, h[0], l[0]);
        split_pair((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w, fmt, h[1], l[1]);
        const long long o = (long long)row * D + 4 * idx;
        *reinterpret_cast<uint2*>(o_hi + o) = make_uint2(h[0], h[1]);
        *reinterpret_cast<uint2*>(o_lo + o) = make_uint2(l[0], l[1]);
      }
    }
  }
}

// ---- multi-head self-attention ----------------------------------------------------------------
// one CTA per (cell, head); thread t owns query row t; K and V of the head live in shared memory.
// two passes over the keys (row max, then exp / sum / PV) keep the softmax in exact fp32.
template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ qkv, int tokens, int heads, int fmt, bf16* __restrict__ o_hi, bf16* __restrict__ o_lo) {
  extern __shared__ float kv[];                   // K [tokens][HD], V [tokens][HD]
  float* Ks = kv;
  float* Vs = kv + tokens * HD;
  const int cell = blockIdx.x / heads, head = blockIdx.x - cell * heads;
  const int D = heads * HD;
  const float* base = qkv + (long long)cell * tokens * 3 * D;
  constexpr int V4 = HD / 4;
  for (int idx = threadIdx.x; idx < tokens * V4; idx += blockDim.x) {
    const int t = idx / V4, d4 = idx - t * V4;
    const float4* rowp = reinterpret_cast<const float4*>(base + (long long)t * 3 * D + head * HD);
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(rowp + (D >> 2) + d4);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(rowp + 2 * (D >> 2) + d4);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= tokens) return;
  const float scale = 1.0f / sqrtf((float)HD);
  float q[HD];
  {
    const float4* qp = reinterpret_cast<const float4*>(base + (long long)t * 3 * D + head * HD);
#pragma unroll
    for (int d4 = 0; d4 < V4; ++d4) {
      const float4 v = __ldg(qp + d4);
      q[4 * d4] = v.x * scale; q[4 * d4 + 1] = v.y * scale; q[4 * d4 + 2] = v.z * scale; q[4 * d4 + 3] = v.w * scale;
    }
  }
  float mx = -INFINITY;
  for (int j = 0; j < tokens; ++j) {
    const float4* kp = reinterpret_cast<const float4*>(Ks + j * HD);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < V4; ++d4) {
      const float4 k = kp[d4];
      s = fmaf(q[4 * d4], k.x, s); s = fmaf(q[4 * d4 + 1], k.y, s); s = fmaf(q[4 * d4 + 2], k.z, s); s = fmaf(q[4 * d4 + 3], k.w, s);
    }
    mx = fmaxf(mx, s);
  }
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float denom = 0.f;
  for (int j = 0; j < tokens; ++j) {
    const float4* kp = reinterpret_cast<const float4*>(Ks + j * HD);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < V4; ++d4) {
      const float4 k = kp[d4];
      s = fmaf(q[4 * d4], k.x, s); s = fmaf(q[4 * d4 + 1], k.y, s); s = fmaf(q[4 * d4 + 2], k.z, s); s = fmaf(q[4 * d4 + 3], k.w, s);
    }
    const float p = expf(s - mx);
    denom += p;
    const float4* vp = reinterpret_cast<const float4*>(Vs + j * HD);
#pragma unroll
    for (int d4 = 0; d4 < V4; ++d4) {
      const float4 v = vp[d4];
      o[4 * d4] = fmaf(p, v.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(p, v.y, o[4 * d4 + 1]);
      o[4 * d4 + 2] = fmaf(p, v.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(p, v.w, o[4 * d4 + 3]);
    }
  }
  const float inv = 1.0f / denom;
  const long long ob = ((long long)cell * tokens + t) * D + head * HD;
#pragma unroll
  for (int d4 = 0; d4 < V4; ++d4) {
    uint32_t h[2], l[2];
    split_pair(o[4 * d4] * inv, o[4 * d4 + 1] * inv, fmt, h[0], l[0]);
    split_pair(o[4 * d4 + 2] * inv, o[4 * d4 + 3] * inv, fmt, h[1], l[1]);
    *reinterpret_cast<uint2*>(o_hi + ob + 4 * d4) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(o_lo + ob + 4 * d4) = make_uint2(l[0], l[1]);
  }
}

// Short sequences (the imputer: 7-16 tokens): several (cell, head) items per 128-thread CTA so that every
// lane has a query row; thread = (item slot, row).  Same arithmetic as attention_kernel.
template <int HD>
__global__ void __launch_bounds__(128)
attention_small_kernel(const float* __restrict__ qkv, int n_items, int tokens, int tpad, int heads, int fmt,
                       bf16* __restrict__ o_hi, bf16* __restrict__ o_lo) {
  extern __shared__ float kv[];                   // per slot: K [tokens][HD], V [tokens][HD]
  const int slots = 128 / tpad;
  const int slot = threadIdx.x / tpad, t = threadIdx.x - slot * tpad;
  const int item = blockIdx.x * slots + slot;
  const int D = heads * HD;
  constexpr int V4 = HD / 4;
  float* Ks = kv + slot * 2 * tokens * HD;
  float* Vs = Ks + tokens * HD;
  const bool live = item < n_items && slot < slots;
  const int cell = live ? item / heads : 0, head = live ? item - cell * heads : 0;
  const float* base = qkv + (long long)cell * tokens * 3 * D;
  if (live) {
    for (int idx = t; idx < tokens * V4; idx += tpad) {
      const int tt = idx / V4, d4 = idx - tt * V4;
      const float4* rowp = reinterpret_cast<const float4*>(base + (long long)tt * 3 * D + head * HD);
      reinterpret_cast<float4*>(Ks)[idx] = __ldg(rowp + (D >> 2) + d4);
      reinterpret_cast<float4*>(Vs)[idx] = __ldg(rowp + 2 * (D >> 2) + d4);
    }
  }
  __syncthreads();
  if (!live || t >= tokens) return;
  const float scale = 1.0f / sqrtf((float)HD);
  float q[HD];
  {
    const float4* qp = reinterpret_cast<const float4*>(base + (long long)t * 3 * D + head * HD);
#pragma unroll
    for (int d4 = 0; d4 < V4; ++d4) {
      const float4 v = __ldg(qp + d4);
      q[4 * d4] = v.x * scale; q[4 * d4 + 1] = v.y * scale; q[4 * d4 + 2] = v.z * scale; q[4 * d4 + 3] = v.w * scale;
    }
  }
  float sc[16];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < tokens) {
      const float4* kp = reinterpret_cast<const float4*>(Ks + j * HD);
      float a = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < V4; ++d4) {
        const float4 k = kp[d4];
        a = fmaf(q[4 * d4], k.x, a); a = fmaf(q[4 * d4 + 1], k.y, a); a = fmaf(q[4 * d4 + 2], k.z, a); a = fmaf(q[4 * d4 + 3], k.w, a);
      }
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
  }
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float denom = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < tokens) {
      const float p = expf(sc[j] - mx);
      denom += p;
      const float4* vp = reinterpret_cast<const float4*>(Vs + j * HD);
#pragma unroll
      for (int d4 = 0; d4 < V4; ++d4) {
        const float4 v = vp[d4];
        o[4 * d4] = fmaf(p, v.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(p, v.y, o[4 * d4 + 1]);
        o[4 * d4 + 2] = fmaf(p, v.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(p, v.w, o[4 * d4 + 3]);
      }
    }
  }
  const float inv = 1.0f / denom;
  const long long ob = ((long long)cell * tokens + t) * D + head * HD;
#pragma unroll
  for (int d4 = 0; d4 < V4; ++d4) {
    uint32_t h[2], l[2];
    split_pair(o[4 * d4] * inv, o[4 * d4 + 1] * inv, fmt, h[0], l[0]);
    split_pair(o[4 * d4 + 2] * inv, o[4 * d4 + 3] * inv, fmt, h[1], l[1]);
    *reinterpret_cast<uint2*>(o_hi + ob + 4 * d4) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(o_lo + ob + 4 * d4) = make_uint2(l[0], l[1]);
  }
}

// ---- class-token rows of the last block ---------------------------------------------------------
// The classifier reads only token 0 of the final LayerNorm (model.py:61-62), and after the last block's attention
// every remaining op (proj, residual, LN2, fc1, GELU, fc2, residual) is row-wise: only the class-token rows are
// computed.  This kernel compacts them: x_cls[cell] = x[cell * tokens], same for the two operand planes of O.
__global__ void __launch_bounds__(256)
gather_cls_rows_kernel(const float* __restrict__ x, const uint16_t* __restrict__ a, long long a_plane, int n_cells, int tokens, int D,
                       float* __restrict__ x_cls, uint16_t* __restrict__ a_cls, long long a_cls_plane) {
  const int per_row = D / 4;                       // float4 / uint2 units
  const long long total = (long long)n_cells * per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long cell = t / per_row;
    const int q = (int)(t - cell * per_row);
    const long long src = cell * tokens * D + 4 * q, dst = cell * D + 4 * q;
    *reinterpret_cast<float4*>(x_cls + dst) = *reinterpret_cast<const float4*>(x + src);
    *reinterpret_cast<uint2*>(a_cls + dst) = *reinterpret_cast<const uint2*>(a + src);
    *reinterpret_cast<uint2*>(a_cls + a_cls_plane + dst) = *reinterpret_cast<const uint2*>(a + a_plane + src);
  }
}

// ---- final LayerNorm on the class token + head + softmax --------------------------------------
// one warp per cell (model.py:61-62 + timm forward_head + softmax(dim=1) of model.py:404)
__global__ void __launch_bounds__(256)
head_softmax_kernel(const float* __restrict__ x, int n_cells, int tokens, int D, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, const float* __restrict__ head_w,
                    const float* __restrict__ head_b, int classes, float* __restrict__ probs, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cell >= n_cells) return;
  const float* xr = x + (long long)cell * tokens * D;
  float v[24];                                   // D <= 768
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const int d = lane + 32 * i; v[i] = d < D ? xr[d] : 0.f; sum += v[i]; }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const int d = lane + 32 * i; if (d < D) { const float a = v[i] - mean; sq += a * a; } }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < 24; ++i) { const int d = lane + 32 * i; v[i] = d < D ? (v[i] - mean) * rstd * __ldg(gamma + d) + __ldg(beta + d) : 0.f; }
  float lg[16];
  for (int k = 0; k < classes; ++k) {
    const float* w = head_w + (long long)k * D;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { const int d = lane + 32 * i; if (d < D) acc = fmaf(v[i], __ldg(w + d), acc); }
    lg[k] = warp_sum(acc) + __ldg(head_b + k);
  }
  if (lane == 0) {
    float mx = lg[0];
    for (int k = 1; k < classes; ++k) mx = fmaxf(mx, lg[k]);
    float e[16], s = 0.f;
    for (int k = 0; k < classes; ++k) { e[k] = expf(lg[k] - mx); s += e[k]; }
    for (int k = 0; k < classes; ++k) {
      probs[(long long)cell * classes + k] = e[k] / s;
      if (logits) logits[(long long)cell * classes + k] = lg[k];
    }
  }
}

int head_softmax_launch(const float* x, int n_cells, int tokens, int D, const float* gamma, const float* beta, float eps,
                        const float* head_w, const float* head_b, int classes, float* probs, float* logits, cudaStream_t st) {
  head_softmax_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(x, n_cells, tokens, D, gamma, beta, eps, head_w, head_b, classes, probs, logits);
  RIBCA_LAUNCH_CHECK("head_softmax_kernel");
  return RIBCA_OK;
}

// ---- MAE glue ---------------------------------------------------------------------------------
struct PresentList { int n; int idx[RIBCA_MAX_PANEL_CH]; int rank[RIBCA_MAX_PANEL_CH]; };

// encoder A operand: row (cell, 0) = 0, row (cell, 1+i) = the 1600 pixels of channel present[i]
__global__ void __launch_bounds__(256)
mae_tiles_split_kernel(const float* __restrict__ patches, int n_cells, int L, const __grid_constant__ PresentList pl,
                       int fmt, bf16* __restrict__ a_hi, bf16* __restrict__ a_lo) {
  const int Te = pl.n + 1;
  const long long total = (long long)n_cells * Te * 400;          // float4 units
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int q = (int)(t % 400);
    const int tok = (int)((t / 400) % Te);
    const long long cell = t / (400ll * Te);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok > 0) v = __ldg(reinterpret_cast<const float4*>(patches + (cell * L + pl.idx[tok - 1]) * 1600ll) + q);
    uint32_t h[2], l[2];
    split_pair(v.x, v.y, fmt, h[0], l[0]);
    split_pair(v.z, v.w, fmt, h[1], l[1]);
    const long long o = (cell * Te + tok) * 1600ll + 4 * q;
    *reinterpret_cast<uint2*>(a_hi + o) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(a_lo + o) = make_uint2(l[0], l[1]);
  }
}

// row table of the encoder patch-embed epilogue: row 0 = cls + pos[0], row 1+i = bias + pos[1+present[i]]
__global__ void mae_enc_table_kernel(const float* __restrict__ cls, const float* __restrict__ bias,
                                     const float* __restrict__ pos, int D, const __grid_constant__ PresentList pl,
                                     float* __restrict__ table) {
  const int Te = pl.n + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Te * D; i += gridDim.x * blockDim.x) {
    const int tok = i / D, d = i - tok * D;
    table[i] = tok == 0 ? cls[d] + pos[d] : bias[d] + pos[(1 + pl.idx[tok - 1]) * D + d];
  }
}

// decoder input (markerImputer.py:208-219): kept tokens back at their positions, mask token elsewhere, + pos
__global__ void __launch_bounds__(256)
mae_decoder_input_kernel(const float* __restrict__ emb, const float* __restrict__ mask_token,
                         const float* __restrict__ dpos, int n_cells, int L, int D,
                         const __grid_constant__ PresentList pl, float* __restrict__ xd) {
  const int Te = pl.n + 1, Td = L + 1;
  const long long total = (long long)n_cells * Td * D;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int d = (int)(t % D);
    const int tok = (int)((t / D) % Td);
    const long long cell = t / ((long long)D * Td);
    float v;
    if (tok == 0) v = emb[(cell * Te) * D + d];
    else {
      const int rk = pl.rank[tok - 1];
      v = rk >= 0 ? emb[(cell * Te + 1 + rk) * D + d] : mask_token[d];
    }
    xd[t] = v + dpos[tok * D + d];
  }
}

// blend of markerImputer.py:312-316: predicted tiles replace the missing channels only
__global__ void __launch_bounds__(256)
mae_scatter_kernel(const float* __restrict__ pred, int n_cells, int L, const __grid_constant__ PresentList pl,
                   float* __restrict__ patches) {
  const int Td = L + 1;
  const long long total = (long long)n_cells * L * 400;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int q = (int)(t % 400);
    const int c = (int)((t / 400) % L);
    const long long cell = t / (400ll * L);
    if (pl.rank[c] >= 0) continue;
    reinterpret_cast<float4*>(patches + (cell * L + c) * 1600ll)[q] =
        reinterpret_cast<const float4*>(pred + (cell * Td + 1 + c) * 1600ll)[q];
  }
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
static int grid_for(long long work_items, int per_block) {
  return (int)std::min<long long>((work_items + per_block - 1) / per_block, (long long)num_sms() * 16);
}

int layernorm_launch(const float* x, int M, int D, const float* g, const float* b, float eps, void* out_split,
                     long long out_plane, int fmt, cudaStream_t st) {
  RIBCA_REQUIRE(D % 4 == 0 && D <= 1024, "layernorm: D=%d must be a multiple of 4 and <= 1024", D);
  if (M <= 0) return RIBCA_OK;
  bf16* hi = static_cast<bf16*>(out_split);
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_LAYERNORM, (double)M * (double)D * 8.0, st);
  // one row per warp, no grid-stride loop: the block scheduler balances the HBM streams better than a capped grid does
  // (4096-cell chunk of vit_l: 0.327 ms with 16 blocks per SM, 0.296 ms with 64, 0.281 ms = 6.8 TB/s uncapped; tools/ln_bench.py)
  const int grid = (M + 7) / 8;
  switch ((D / 4 + 31) / 32) {
#define RIBCA_LN_CASE(NI) case NI: layernorm_split_kernel<NI><<<grid, 256, 0, st>>>(x, M, D, g, b, eps, fmt, hi, hi + out_plane); break;
    RIBCA_LN_CASE(1) RIBCA_LN_CASE(2) RIBCA_LN_CASE(3) RIBCA_LN_CASE(4) RIBCA_LN_CASE(5) RIBCA_LN_CASE(6) RIBCA_LN_CASE(7)
    default: layernorm_split_kernel<8><<<grid, 256, 0, st>>>(x, M, D, g, b, eps, fmt, hi, hi + out_plane); break;
#undef RIBCA_LN_CASE
  }
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("layernorm_split_kernel");
  return RIBCA_OK;
}

template <int HD>
static int attention_launch_hd(const float* qkv, int cells, int tokens, int heads, int fmt, bf16* hi, bf16* lo, cudaStream_t st) {
  if (tokens <= 16) {       // packed short-sequence kernel
    const int tpad = tokens <= 8 ? 8 : 16, slots = 128 / tpad, n_items = cells * heads;
    const size_t smem_small = (size_t)slots * 2 * tokens * HD * sizeof(float);
    RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_small_kernel<HD>), (int)(16 * 2 * 16 * HD * 4), "cudaFuncSetAttribute(attention_small_kernel)"));
    const bool prof_s = profiling();
    if (prof_s) prof_begin_span(RIBCA_PROF_ATTENTION, 4.0 * (double)cells * heads * (double)tokens * tokens * HD, st);
    attention_small_kernel<HD><<<(n_items + slots - 1) / slots, 128, smem_small, st>>>(qkv, n_items, tokens, tpad, heads, fmt, hi, lo);
    if (prof_s) prof_end_span(st);
    RIBCA_LAUNCH_CHECK("attention_small_kernel");
    return RIBCA_OK;
  }
  const size_t smem = (size_t)2 * tokens * HD * sizeof(float);
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_kernel<HD>), (int)(2 * 128 * HD * 4), "cudaFuncSetAttribute(attention_kernel)"));
  const int threads = (tokens + 31) / 32 * 32;
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_ATTENTION, 4.0 * (double)cells * heads * (double)tokens * tokens * HD, st);
  attention_kernel<HD><<<cells * heads, threads, smem, st>>>(qkv, tokens, heads, fmt, hi, lo);
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("attention_kernel");
  return RIBCA_OK;
}

int attention_launch(const float* qkv, int cells, int tokens, int heads, int hd, void* out_split, long long out_plane,
                     int fmt, cudaStream_t st) {
  RIBCA_REQUIRE(tokens > 0 && tokens <= 128, "attention: tokens=%d outside [1,128]", tokens);
  if (cells <= 0) return RIBCA_OK;
  bf16* hi = static_cast<bf16*>(out_split);
  bf16* lo = hi + out_plane;
  switch (hd) {
    case 12: return attention_launch_hd<12>(qkv, cells, tokens, heads, fmt, hi, lo, st);
    case 24: return attention_launch_hd<24>(qkv, cells, tokens, heads, fmt, hi, lo, st);
    case 32: return attention_launch_hd<32>(qkv, cells, tokens, heads, fmt, hi, lo, st);
    case 48: return attention_launch_hd<48>(qkv, cells, tokens, heads, fmt, hi, lo, st);
    case 64: return attention_launch_hd<64>(qkv, cells, tokens, heads, fmt, hi, lo, st);
    default: set_error("attention: unsupported head_dim %d", hd); return RIBCA_EUNSUPPORTED;
  }
}

constexpr int kInterleaveMinCells = 512;   // below this the GEMMs are too short for a second stream to pay

struct BlockBuffers {
  float* x;          // [M][D]
  bf16* a;           // split [2][M][D]: LayerNorm output, then the attention output
  float* qkv;        // [M][3D]
  bf16* h;           // split [2][M][4D]
  bf16* xa;          // LayerNorm-folded flow: operand planes [2][M][D] of the RAW residual stream x ...
  float* stats;      // ... and its per-row partial sums [M][RIBCA_LN_SLOTS][2] (both written by the *_LN epilogues)
};

// timm Block x depth: x += proj(attn(LN1 x)); x += fc2(gelu(fc1(LN2 x)))
// cls_rows != nullptr: the last block computes proj / MLP for the class-token rows only and leaves them,
// compacted, in *cls_rows ([cells][D] fp32, carved from the MLP buffer)
// folded: no LayerNorm kernel.  The GEMM that writes x (patch embedding before the first block, proj, fc2: *_LN epilogues)
// also leaves the operand planes of the raw x in b.xa and its row statistics in b.stats; qkv and fc1 contract those planes
// with W * diag(gamma) and apply mean / rstd / beta in their epilogue (include/ribca_b200.h: ribca_ln_fold).
static int run_blocks(const ribca_block_desc* blocks, int depth, int D, int heads, int cells, int tokens,
                      const float* wf32, const bf16* wsplit, long long split_plane, const BlockBuffers& b,
                      int precision, int wls, cudaStream_t st, float** cls_rows = nullptr, int fold_mask = 0) {
  const bool fold1 = (fold_mask & 1) != 0;      // norm1 inside the qkv GEMM (producers: patch embedding, fc2)
  const bool fold2 = (fold_mask & 2) != 0;      // norm2 inside the fc1 GEMM (producer: proj)
  const bool folded = fold_mask != 0;
  const int M = cells * tokens;
  const int fmt = fmt_of(precision);
  const long long pa = (long long)M * D, ph = (long long)M * 4 * D;
  const int hd = D / heads, hdp = (hd + 15) / 16 * 16;
  const int Wq = 3 * heads * hdp;                       // head-padded qkv width
  const bool tensor_attention = tokens > 32 && tokens <= 112;   // tcgen05 for the classifiers, FP32 pipe for the imputer
  RIBCA_REQUIRE(tensor_attention || hdp == hd, "short-sequence attention needs head_dim %% 16 == 0");
  RIBCA_REQUIRE(!folded || tensor_attention, "the LayerNorm-folded flow is the classifiers' (tensor-core attention)");
  const int slots = folded ? gemm_ln_slots(D, precision) : 0;      // every producer of x has N = D
  ribca_ln_fold ln_in{b.stats, nullptr, slots, 1e-6f, nullptr};  // consumer side (c1 set per GEMM)
  const ribca_ln_fold ln_out{nullptr, nullptr, 0, 1e-6f, b.stats};
  for (int l = 0; l < depth; ++l) {
    const ribca_block_desc& w = blocks[l];
    if (!fold1) RIBCA_TRY(layernorm_launch(b.x, M, D, wf32 + w.ln1_g, wf32 + w.ln1_b, 1e-6f, b.a, pa, fmt, st));
    if (tensor_attention) {
      bf16* qs = reinterpret_cast<bf16*>(b.qkv);
      const long long pq = (long long)M * Wq;
      if (fold1) {
        ln_in.c1 = wf32 + w.qkv_c1;
        RIBCA_TRY(gemm_launch(b.xa, pa, wsplit + w.qkv_w, split_plane, M, Wq, D, wf32 + w.qkv_c2, nullptr, 0,
                              RIBCA_EPI_STORE_SPLIT, nullptr, qs, pq, precision, wls, st, &ln_in));
      } else {
        RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.qkv_w, split_plane, M, Wq, D, wf32 + w.qkv_b, nullptr, 0,
                              RIBCA_EPI_STORE_SPLIT, nullptr, qs, pq, precision, wls, st));
      }
      RIBCA_TRY(attention_tc_launch(qs, pq, cells, tokens, heads, hd, b.a, pa, fmt, st));
    } else {
      RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.qkv_w, split_plane, M, Wq, D, wf32 + w.qkv_b, nullptr, 0,
                            RIBCA_EPI_STORE, b.qkv, nullptr, 0, precision, wls, st));
      RIBCA_TRY(attention_launch(b.qkv, cells, tokens, heads, hd, b.a, pa, fmt, st));
    }
    if (cls_rows && l == depth - 1) {
      // compact buffers inside the (idle) MLP buffer: x_cls fp32 [cells][D], a_cls planes [2][cells][D], h_cls planes [2][cells][4D]
      // (+ folded: xa_cls planes [2][cells][D]; the statistics of the compact rows reuse the head of b.stats)
      char* base = reinterpret_cast<char*>(b.h);
      float* x_cls = reinterpret_cast<float*>(base);
      bf16* a_cls = reinterpret_cast<bf16*>(base + align_up((size_t)cells * D * 4, 256));
      bf16* h_cls = reinterpret_cast<bf16*>(reinterpret_cast<char*>(a_cls) + align_up((size_t)cells * D * 4, 256));
      bf16* xa_cls = reinterpret_cast<bf16*>(reinterpret_cast<char*>(h_cls) + align_up((size_t)cells * D * 16, 256));
      const long long pc = (long long)cells * D, phc = (long long)cells * 4 * D;
      gather_cls_rows_kernel<<<grid_for((long long)cells * (D / 4), 256), 256, 0, st>>>(
          b.x, reinterpret_cast<const uint16_t*>(b.a), pa, cells, tokens, D, x_cls, reinterpret_cast<uint16_t*>(a_cls), pc);
      RIBCA_LAUNCH_CHECK("gather_cls_rows_kernel");
      if (fold2) {
        RIBCA_TRY(gemm_launch(a_cls, pc, wsplit + w.proj_w, split_plane, cells, D, D, wf32 + w.proj_b, nullptr, 0,
                              RIBCA_EPI_RESIDUAL_LN, x_cls, xa_cls, pc, precision, wls, st, &ln_out));
        ln_in.c1 = wf32 + w.fc1_c1;
        RIBCA_TRY(gemm_launch(xa_cls, pc, wsplit + w.fc1_w, split_plane, cells, 4 * D, D, wf32 + w.fc1_c2, nullptr, 0,
                              RIBCA_EPI_GELU, nullptr, h_cls, phc, precision, wls, st, &ln_in));
      } else {
        RIBCA_TRY(gemm_launch(a_cls, pc, wsplit + w.proj_w, split_plane, cells, D, D, wf32 + w.proj_b, nullptr, 0,
                              RIBCA_EPI_RESIDUAL, x_cls, nullptr, 0, precision, wls, st));
        RIBCA_TRY(layernorm_launch(x_cls, cells, D, wf32 + w.ln2_g, wf32 + w.ln2_b, 1e-6f, a_cls, pc, fmt, st));
        RIBCA_TRY(gemm_launch(a_cls, pc, wsplit + w.fc1_w, split_plane, cells, 4 * D, D, wf32 + w.fc1_b, nullptr, 0,
                              RIBCA_EPI_GELU, nullptr, h_cls, phc, precision, wls, st));
      }
      RIBCA_TRY(gemm_launch(h_cls, phc, wsplit + w.fc2_w, split_plane, cells, D, 4 * D, wf32 + w.fc2_b, nullptr, 0,
                            RIBCA_EPI_RESIDUAL, x_cls, nullptr, 0, precision, wls, st));
      *cls_rows = x_cls;
      return RIBCA_OK;
    }
    if (folded) {
      if (fold2) {
        RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.proj_w, split_plane, M, D, D, wf32 + w.proj_b, nullptr, 0,
                              RIBCA_EPI_RESIDUAL_LN, b.x, b.xa, pa, precision, wls, st, &ln_out));
        ln_in.c1 = wf32 + w.fc1_c1;
        RIBCA_TRY(gemm_launch(b.xa, pa, wsplit + w.fc1_w, split_plane, M, 4 * D, D, wf32 + w.fc1_c2, nullptr, 0,
                              RIBCA_EPI_GELU, nullptr, b.h, ph, precision, wls, st, &ln_in));
      } else {
        RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.proj_w, split_plane, M, D, D, wf32 + w.proj_b, nullptr, 0,
                              RIBCA_EPI_RESIDUAL, b.x, nullptr, 0, precision, wls, st));
        RIBCA_TRY(layernorm_launch(b.x, M, D, wf32 + w.ln2_g, wf32 + w.ln2_b, 1e-6f, b.a, pa, fmt, st));
        RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.fc1_w, split_plane, M, 4 * D, D, wf32 + w.fc1_b, nullptr, 0,
                              RIBCA_EPI_GELU, nullptr, b.h, ph, precision, wls, st));
      }
      // fc2 leaves the next block's norm1 inputs (planes + statistics) when norm1 is folded
      RIBCA_TRY(gemm_launch(b.h, ph, wsplit + w.fc2_w, split_plane, M, D, 4 * D, wf32 + w.fc2_b, nullptr, 0,
                            fold1 ? RIBCA_EPI_RESIDUAL_LN : RIBCA_EPI_RESIDUAL, b.x, fold1 ? b.xa : nullptr, fold1 ? pa : 0, precision, wls, st,
                            fold1 ? &ln_out : nullptr));
      continue;
    }
    RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.proj_w, split_plane, M, D, D, wf32 + w.proj_b, nullptr, 0,
                          RIBCA_EPI_RESIDUAL, b.x, nullptr, 0, precision, wls, st));
    RIBCA_TRY(layernorm_launch(b.x, M, D, wf32 + w.ln2_g, wf32 + w.ln2_b, 1e-6f, b.a, pa, fmt, st));
    RIBCA_TRY(gemm_launch(b.a, pa, wsplit + w.fc1_w, split_plane, M, 4 * D, D, wf32 + w.fc1_b, nullptr, 0,
                          RIBCA_EPI_GELU, nullptr, b.h, ph, precision, wls, st));
    RIBCA_TRY(gemm_launch(b.h, ph, wsplit + w.fc2_w, split_plane, M, D, 4 * D, wf32 + w.fc2_b, nullptr, 0,
                          RIBCA_EPI_RESIDUAL, b.x, nullptr, 0, precision, wls, st));
  }
  return RIBCA_OK;
}

struct Carver {
  char* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

static size_t block_buffers(Carver& cv, BlockBuffers& b, long long M, int D, int heads, bool folded = false) {
  const int hdp = (D / heads + 15) / 16 * 16;
  b.x = cv.take<float>(M * D);
  b.a = cv.take<bf16>(2 * M * D);
  b.qkv = cv.take<float>(M * 3 * heads * hdp);     // fp32 [M][3D] or split-bf16 [2][M][3*heads*hdp]: same bytes
  b.h = cv.take<bf16>(2 * M * 4 * D);
  b.xa = folded ? cv.take<bf16>(2 * M * D) : nullptr;
  b.stats = folded ? cv.take<float>(M * RIBCA_LN_SLOTS * 2) : nullptr;
  return cv.off;
}

}  // namespace ribca

using namespace ribca;

extern "C" {

int ribca_layernorm_split(const float* x, int M, int D, const float* gamma, const float* beta, float eps,
                          void* out_split, long long out_plane, int format, ribca_stream_t stream) {
  RIBCA_REQUIRE(x && gamma && beta && out_split, "ribca_layernorm_split: null pointer");
  RIBCA_REQUIRE(format == kFmtBf16 || format == kFmtF16F8, "ribca_layernorm_split: unknown plane format %d", format);
  return layernorm_launch(x, M, D, gamma, beta, eps, out_split, out_plane, format, as_stream(stream));
}

int ribca_attention(const float* qkv, int cells, int tokens, int heads, int head_dim, void* out_split,
                    long long out_plane, int format, ribca_stream_t stream) {
  RIBCA_REQUIRE(qkv && out_split && heads > 0, "ribca_attention: bad arguments");
  RIBCA_REQUIRE(format == kFmtBf16 || format == kFmtF16F8, "ribca_attention: unknown plane format %d", format);
  return attention_launch(qkv, cells, tokens, heads, head_dim, out_split, out_plane, format, as_stream(stream));
}

static int interleave_split(int n_cells);

// the larger of the serial carve-up and the two-halves carve-up (they differ by alignment padding only)
size_t ribca_vit_workspace_bytes(const ribca_vit_desc* desc, int n_cells) {
  if (!desc || n_cells <= 0) return 0;
  Carver cv{nullptr, 0};
  BlockBuffers b;
  const bool folded = desc->ln_folded != 0;
  block_buffers(cv, b, (long long)n_cells * desc->tokens, desc->dim, desc->heads, folded);
  size_t need = cv.off;
  const int n0 = interleave_split(n_cells);
  if (n0 > 0) {
    Carver c2{nullptr, 0};
    block_buffers(c2, b, (long long)n0 * desc->tokens, desc->dim, desc->heads, folded);
    block_buffers(c2, b, (long long)(n_cells - n0) * desc->tokens, desc->dim, desc->heads, folded);
    need = std::max(need, c2.off);
  }
  return align_up(need, 256);
}

// one contiguous range of cells through the classifier on one stream (buffers carved by the caller)
static int vit_forward_range(const ribca_vit_desc* desc, const float* wf32, const bf16* wsplit, const float* patches,
                             int n_cells, float* probs, float* logits, const BlockBuffers& b, int precision, cudaStream_t st) {
  const int D = desc->dim, T = desc->tokens, C = desc->in_chans;
  const long long M = (long long)n_cells * T;
  // patch embedding: im2col into the (larger) MLP buffer, GEMM with the cls/pos/bias row table
  const int Kpe = 16 * C;
  const long long pe_plane = M * Kpe;
  const int fmt = fmt_of(precision), wls = desc->w_log2_scale;
  im2col_split_kernel<<<grid_for((long long)n_cells * C * 400, 256), 256, 0, st>>>(patches, n_cells, C, fmt, b.h, b.h + pe_plane);
  RIBCA_LAUNCH_CHECK("im2col_split_kernel");
  const int folded = desc->ln_folded;
  if (folded & 1) {
    const ribca_ln_fold ln_out{nullptr, nullptr, 0, 1e-6f, b.stats};
    RIBCA_TRY(gemm_launch(b.h, pe_plane, wsplit + desc->embed_w, desc->split_plane, (int)M, D, Kpe, nullptr,
                          wf32 + desc->embed_table, T, RIBCA_EPI_STORE_LN, b.x, b.xa, M * D, precision, wls, st, &ln_out));
  } else {
    RIBCA_TRY(gemm_launch(b.h, pe_plane, wsplit + desc->embed_w, desc->split_plane, (int)M, D, Kpe, nullptr,
                          wf32 + desc->embed_table, T, RIBCA_EPI_STORE, b.x, nullptr, 0, precision, wls, st));
  }
  float* x_cls = nullptr;
  RIBCA_TRY(run_blocks(desc->blocks, desc->depth, D, desc->heads, n_cells, T, wf32, wsplit, desc->split_plane, b, precision, wls, st, &x_cls, folded));
  head_softmax_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(x_cls, n_cells, 1, D, wf32 + desc->norm_g, wf32 + desc->norm_b, 1e-6f,
                                                         wf32 + desc->head_w, wf32 + desc->head_b, desc->classes, probs, logits);
  RIBCA_LAUNCH_CHECK("head_softmax_kernel");
  return RIBCA_OK;
}

// cells of the first half of a two-way interleaved call (0 = run the call as one range)
static int interleave_split(int n_cells) { return n_cells >= kInterleaveMinCells ? (n_cells + 1) / 2 : 0; }

int ribca_vit_forward(const ribca_vit_desc* desc, const float* wf32, const void* wsplit_, const float* patches,
                      int n_cells, float* probs, float* logits, void* workspace, size_t workspace_bytes,
                      int precision, ribca_stream_t stream) {
  RIBCA_REQUIRE(desc && wf32 && wsplit_ && patches && probs && workspace, "ribca_vit_forward: null pointer");
  RIBCA_REQUIRE(desc->tokens == 101 && desc->depth > 0 && desc->depth <= 16 && desc->classes <= 16 && desc->dim <= 768 &&
                    desc->dim % desc->heads == 0 && desc->in_chans <= RIBCA_MAX_PANEL_CH,
                "ribca_vit_forward: unsupported model shape");
  if (n_cells <= 0) return RIBCA_OK;
  if (workspace_bytes < ribca_vit_workspace_bytes(desc, n_cells)) {
    set_error("ribca_vit_forward: workspace %zu < %zu", workspace_bytes, ribca_vit_workspace_bytes(desc, n_cells));
    return RIBCA_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const bf16* wsplit = static_cast<const bf16*>(wsplit_);
  const int D = desc->dim, T = desc->tokens;
  const long long M = (long long)n_cells * T;
  RIBCA_REQUIRE(M < (1ll << 31) / 16, "ribca_vit_forward: %d cells per call is too many; chunk the batch", n_cells);
  Carver cv{static_cast<char*>(workspace), 0};
  if (precision == RIBCA_FP32) {          // the reference's own arithmetic on the FP32 pipe (stage4_fp32.cu)
    RIBCA_REQUIRE(desc->plane_format == RIBCA_PLANES_F32, "ribca_vit_forward: RIBCA_FP32 needs fp32 weight matrices (plane format %d given)",
                  desc->plane_format);
    BlockBuffers b;
    block_buffers(cv, b, M, D, desc->heads);
    return vit_forward_f32(desc, wf32, static_cast<const float*>(wsplit_), patches, n_cells, probs, logits, b.x,
                           reinterpret_cast<float*>(b.a), b.qkv, reinterpret_cast<float*>(b.h), st);
  }
  RIBCA_REQUIRE(desc->ln_folded >= 0 && desc->ln_folded <= 3 && (!desc->ln_folded || desc->dim % 16 == 0),
                "ribca_vit_forward: LayerNorm fold mask %d needs dim %% 16 == 0 (dim %d)", desc->ln_folded, desc->dim);
  RIBCA_REQUIRE(desc->plane_format == fmt_of(precision), "ribca_vit_forward: weights are packed in plane format %d but precision %d needs %d",
                desc->plane_format, precision, fmt_of(precision));
  const int n0 = interleave_enabled() ? interleave_split(n_cells) : 0;
  const bool folded = desc->ln_folded != 0;
  if (n0 == 0) {
    BlockBuffers b;
    block_buffers(cv, b, M, D, desc->heads, folded);
    return vit_forward_range(desc, wf32, wsplit, patches, n_cells, probs, logits, b, precision, st);
  }
  // two-way interleave: half 0 on the caller's stream, half 1 on the side stream (include/ribca_b200.h: ribca_set_interleave)
  const int n1 = n_cells - n0;
  BlockBuffers b0, b1;
  block_buffers(cv, b0, (long long)n0 * T, D, desc->heads, folded);
  block_buffers(cv, b1, (long long)n1 * T, D, desc->heads, folded);
  SideStream side;
  RIBCA_TRY(side_stream(&side));
  RIBCA_TRY(check_cuda(cudaEventRecord(side.fork, st), "cudaEventRecord(fork)"));
  RIBCA_TRY(check_cuda(cudaStreamWaitEvent(side.stream, side.fork, 0), "cudaStreamWaitEvent(fork)"));
  const long long per_cell = (long long)desc->in_chans * 1600;
  const int rc1 = vit_forward_range(desc, wf32, wsplit, patches + n0 * per_cell, n1, probs + (long long)n0 * desc->classes,
                                    logits ? logits + (long long)n0 * desc->classes : nullptr, b1, precision, side.stream);
  const int rc0 = vit_forward_range(desc, wf32, wsplit, patches, n0, probs, logits, b0, precision, st);
  // always join, also after a failed launch: the caller's stream must not run ahead of the side stream
  const int rcj = check_cuda(cudaEventRecord(side.join, side.stream), "cudaEventRecord(join)");
  const int rcw = rcj == RIBCA_OK ? check_cuda(cudaStreamWaitEvent(st, side.join, 0), "cudaStreamWaitEvent(join)") : rcj;
  return rc1 != RIBCA_OK ? rc1 : rc0 != RIBCA_OK ? rc0 : rcw;
}

static void mae_carve(const ribca_mae_desc* d, int n_cells, int n_present, Carver& cv, BlockBuffers& be, BlockBuffers& bd,
                      float*& emb, float*& pred, float*& table) {
  const long long Me = (long long)n_cells * (n_present + 1), Md = (long long)n_cells * (d->channels + 1);
  block_buffers(cv, be, Me, d->enc_dim, d->enc_heads);
  block_buffers(cv, bd, Md, d->dec_dim, d->dec_heads);
  emb = cv.take<float>(Me * d->dec_dim);
  pred = cv.take<float>(Md * 1600);
  table = cv.take<float>((size_t)(n_present + 1) * d->enc_dim);
}

size_t ribca_mae_workspace_bytes(const ribca_mae_desc* desc, int n_cells) {
  if (!desc || n_cells <= 0) return 0;
  Carver cv{nullptr, 0};
  BlockBuffers be, bd; float *emb, *pred, *table;
  mae_carve(desc, n_cells, desc->channels, cv, be, bd, emb, pred, table);   // worst case: every marker present
  return align_up(cv.off, 256);
}

int ribca_mae_impute(const ribca_mae_desc* desc, const float* wf32, const void* wsplit_, float* patches, int n_cells,
                     const int* h_present, int n_present, void* workspace, size_t workspace_bytes, int precision,
                     ribca_stream_t stream) {
  RIBCA_REQUIRE(desc && wf32 && wsplit_ && patches && h_present && workspace, "ribca_mae_impute: null pointer");
  const int L = desc->channels;
  RIBCA_REQUIRE(L > 0 && L <= RIBCA_MAX_PANEL_CH && n_present > 0 && n_present <= L, "ribca_mae_impute: bad channel counts");
  RIBCA_REQUIRE(desc->enc_depth <= 16 && desc->dec_depth <= 16, "ribca_mae_impute: too many blocks");
  if (n_cells <= 0 || n_present == L) return RIBCA_OK;
  if (workspace_bytes < ribca_mae_workspace_bytes(desc, n_cells)) {
    set_error("ribca_mae_impute: workspace %zu < %zu", workspace_bytes, ribca_mae_workspace_bytes(desc, n_cells));
    return RIBCA_EWORKSPACE;
  }
  PresentList pl;
  memset(&pl, 0, sizeof(pl));
  pl.n = n_present;
  for (int c = 0; c < RIBCA_MAX_PANEL_CH; ++c) pl.rank[c] = -1;
  for (int i = 0; i < n_present; ++i) {
    RIBCA_REQUIRE(h_present[i] >= 0 && h_present[i] < L && (i == 0 || h_present[i] > h_present[i - 1]),
                  "ribca_mae_impute: present list must be ascending positions in [0,%d)", L);
    pl.idx[i] = h_present[i];
    pl.rank[h_present[i]] = i;
  }
  cudaStream_t st = as_stream(stream);
  const bf16* wsplit = static_cast<const bf16*>(wsplit_);
  const int De = desc->enc_dim, Dd = desc->dec_dim, Te = n_present + 1, Td = L + 1;
  const long long Me = (long long)n_cells * Te, Md = (long long)n_cells * Td;
  Carver cv{static_cast<char*>(workspace), 0};
  BlockBuffers be, bd; float *emb, *pred, *table;
  mae_carve(desc, n_cells, n_present, cv, be, bd, emb, pred, table);

  // encoder (markerImputer.py:186-206)
  const long long tile_plane = Me * 1600;
  const int fmt = fmt_of(precision), wls = desc->w_log2_scale;
  RIBCA_REQUIRE(desc->plane_format == fmt, "ribca_mae_impute: weights are packed in plane format %d but precision %d needs %d",
                desc->plane_format, precision, fmt);
  mae_tiles_split_kernel<<<grid_for(Me * 400, 256), 256, 0, st>>>(patches, n_cells, L, pl, fmt, be.h, be.h + tile_plane);
  RIBCA_LAUNCH_CHECK("mae_tiles_split_kernel");
  mae_enc_table_kernel<<<grid_for((long long)Te * De, 256), 256, 0, st>>>(wf32 + desc->cls_token, wf32 + desc->embed_bias,
                                                                         wf32 + desc->pos_embed, De, pl, table);
  RIBCA_LAUNCH_CHECK("mae_enc_table_kernel");
  RIBCA_TRY(gemm_launch(be.h, tile_plane, wsplit + desc->embed_w, desc->split_plane, (int)Me, De, 1600, nullptr, table, Te,
                        RIBCA_EPI_STORE, be.x, nullptr, 0, precision, wls, st));
  RIBCA_TRY(run_blocks(desc->enc_blocks, desc->enc_depth, De, desc->enc_heads, n_cells, Te, wf32, wsplit, desc->split_plane, be, precision, wls, st));
  RIBCA_TRY(layernorm_launch(be.x, (int)Me, De, wf32 + desc->norm_g, wf32 + desc->norm_b, 1e-6f, be.a, Me * De, fmt, st));
  // decoder (markerImputer.py:208-232)
  RIBCA_TRY(gemm_launch(be.a, Me * De, wsplit + desc->dec_embed_w, desc->split_plane, (int)Me, Dd, De, wf32 + desc->dec_embed_b,
                        nullptr, 0, RIBCA_EPI_STORE, emb, nullptr, 0, precision, wls, st));
  mae_decoder_input_kernel<<<grid_for(Md * Dd, 256), 256, 0, st>>>(emb, wf32 + desc->mask_token, wf32 + desc->dec_pos_embed,
                                                                    n_cells, L, Dd, pl, bd.x);
  RIBCA_LAUNCH_CHECK("mae_decoder_input_kernel");
  RIBCA_TRY(run_blocks(desc->dec_blocks, desc->dec_depth, Dd, desc->dec_heads, n_cells, Td, wf32, wsplit, desc->split_plane, bd, precision, wls, st));
  RIBCA_TRY(layernorm_launch(bd.x, (int)Md, Dd, wf32 + desc->dec_norm_g, wf32 + desc->dec_norm_b, 1e-6f, bd.a, Md * Dd, fmt, st));
  RIBCA_TRY(gemm_launch(bd.a, Md * Dd, wsplit + desc->pred_w, desc->split_plane, (int)Md, 1600, Dd, wf32 + desc->pred_b, nullptr, 0,
                        RIBCA_EPI_STORE, pred, nullptr, 0, precision, wls, st));
  mae_scatter_kernel<<<grid_for((long long)n_cells * L * 400, 256), 256, 0, st>>>(pred, n_cells, L, pl, patches);
  RIBCA_LAUNCH_CHECK("mae_scatter_kernel");
  return RIBCA_OK;
}

}  // extern "C"
