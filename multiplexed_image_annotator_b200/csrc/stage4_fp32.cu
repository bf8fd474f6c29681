// Stage 4 in plain fp32 on the FP32 pipe (precision RIBCA_FP32): the arithmetic of the reference itself
// (cta/model.py:397-406 runs its ViT in fp32, cta/markerImputer.py:308-317 its MAE) - fp32 operands, fp32 FMA
// accumulation, erff GELU, expf softmax.  It is the LAST level of the exact-label re-evaluation
// (pipeline.refine_labels): only the few cells whose decision margin is still below the error bound of the
// split-bf16 tensor-core pass come here, so the kernels are sized for correctness and a fair FFMA rate
// (128 x 128 x 16 register-blocked SGEMM), not for the tensor cores.
#include "common.cuh"

namespace ribca {

int head_softmax_launch(const float* x, int n_cells, int tokens, int D, const float* gamma, const float* beta, float eps,
                        const float* head_w, const float* head_b, int classes, float* probs, float* logits, cudaStream_t st);

constexpr int FT = 128;     // output tile edge
constexpr int FK = 16;      // K step
constexpr int FPAD = 4;

struct SgemmEpi {
  const float* bias;        // [N] or null
  const float* row_table;   // [period][N] or null
  int table_period;
  int mode;                 // RIBCA_EPI_STORE / RIBCA_EPI_RESIDUAL / RIBCA_EPI_GELU (fp32 out in every mode)
  float* out;               // [M][N]
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// C[M,N] = A[M,K] . W[N,K]^T; A, W row-major with K contiguous (PyTorch Linear layout); K % 4 == 0, N % 4 == 0
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ W, int M, int N, int K, const SgemmEpi epi) {
  __shared__ float As[2][FK][FT + FPAD];
  __shared__ float Ws[2][FK][FT + FPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * FT, n0 = blockIdx.x * FT;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float4 ra[2], rw[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 2, kq = idx & 3;
      const int k = k0 + kq * 4;
      ra[i] = (m0 + r < M && k < K) ? __ldg(reinterpret_cast<const float4*>(A + (long long)(m0 + r) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[i] = (n0 + r < N && k < K) ? __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + r) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      Ws[buf][kq][r] = rw[i].x; Ws[buf][kq + 1][r] = rw[i].y; Ws[buf][kq + 2][r] = rw[i].z; Ws[buf][kq + 3][r] = rw[i].w;
    }
  };
  const int n_steps = (K + FK - 1) / FK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int s = 0; s < n_steps; ++s) {
    const int buf = s & 1;
    if (s + 1 < n_steps) fetch((s + 1) * FK);
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (s + 1 < n_steps) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
    const float* trow = epi.row_table ? epi.row_table + (long long)(row % epi.table_period) * N : nullptr;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= N) continue;
      float v[4] = {acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (epi.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(epi.bias + col));
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      if (trow) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(trow + col));
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      float4* o = reinterpret_cast<float4*>(epi.out + (long long)row * N + col);
      if (epi.mode == RIBCA_EPI_GELU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = gelu_exact(v[q]);
      } else if (epi.mode == RIBCA_EPI_RESIDUAL) {
        const float4 x = *o;
        v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
      }
      *o = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

static int sgemm(const float* A, const float* W, int M, int N, int K, const float* bias, const float* table, int period, int mode,
                 float* out, cudaStream_t st) {
  RIBCA_REQUIRE(M > 0 && N > 0 && K > 0 && N % 4 == 0 && K % 4 == 0, "sgemm: bad shape M=%d N=%d K=%d", M, N, K);
  SgemmEpi epi{bias, table, period > 0 ? period : 1, mode, out};
  dim3 grid((N + FT - 1) / FT, (M + FT - 1) / FT);
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_GEMM_F32, 2.0 * (double)M * (double)N * (double)K, st);
  sgemm_nt_kernel<<<grid, 256, 0, st>>>(A, W, M, N, K, epi);
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("sgemm_nt_kernel");
  return RIBCA_OK;
}

// LayerNorm, fp32 in / fp32 out, one warp per row (two-pass statistics as torch.nn.LayerNorm)
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, int M, int D, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    const float* xr = x + (long long)row * D;
    float sum = 0.f;
    for (int d = lane; d < D; d += 32) sum += xr[d];
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
    for (int d = lane; d < D; d += 32) { const float a = xr[d] - mean; sq += a * a; }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + eps);
    float* o = out + (long long)row * D;
    for (int d = lane; d < D; d += 32) o[d] = (xr[d] - mean) * rstd * __ldg(gamma + d) + __ldg(beta + d);
  }
}

// softmax(q k^T / sqrt(hd)) v for one (cell, head) per CTA, thread = query row; qkv fp32 [M][3 * heads * HDP]
// (head-padded like the packed weights, padding columns are exact zeros), output fp32 [M][heads * hd]
template <int HDP>
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ qkv, int tokens, int heads, int hd, float* __restrict__ out) {
  extern __shared__ float kv[];
  float* Ks = kv;
  float* Vs = kv + tokens * HDP;
  const int cell = blockIdx.x / heads, head = blockIdx.x - cell * heads;
  const int Wq = 3 * heads * HDP;
  const float* base = qkv + (long long)cell * tokens * Wq;
  constexpr int V4 = HDP / 4;
  for (int idx = threadIdx.x; idx < tokens * V4; idx += blockDim.x) {
    const int t = idx / V4, d4 = idx - t * V4;
    const float4* rowp = reinterpret_cast<const float4*>(base + (long long)t * Wq + head * HDP);
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(rowp + (heads * HDP >> 2) + d4);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(rowp + 2 * (heads * HDP >> 2) + d4);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= tokens) return;
  const float scale = 1.0f / sqrtf((float)hd);
  float q[HDP];
  {
    const float4* qp = reinterpret_cast<const float4*>(base + (long long)t * Wq + head * HDP);
#pragma unroll
    for (int d4 = 0; d4 < V4; ++d4) {
      const float4 v = __ldg(qp + d4);
      q[4 * d4] = v.x * scale; q[4 * d4 + 1] = v.y * scale; q[4 * d4 + 2] = v.z * scale; q[4 * d4 + 3] = v.w * scale;
    }
  }
  float mx = -INFINITY;
  for (int j = 0; j < tokens; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HDP; ++d) s = fmaf(q[d], Ks[j * HDP + d], s);
    mx = fmaxf(mx, s);
  }
  float o[HDP];
#pragma unroll
  for (int d = 0; d < HDP; ++d) o[d] = 0.f;
  float denom = 0.f;
  for (int j = 0; j < tokens; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HDP; ++d) s = fmaf(q[d], Ks[j * HDP + d], s);
    const float p = expf(s - mx);
    denom += p;
#pragma unroll
    for (int d = 0; d < HDP; ++d) o[d] = fmaf(p, Vs[j * HDP + d], o[d]);
  }
  const float inv = 1.0f / denom;
  float* orow = out + ((long long)cell * tokens + t) * (heads * hd) + head * hd;
#pragma unroll
  for (int d = 0; d < HDP; ++d)
    if (d < hd) orow[d] = o[d] * inv;
}

template <int HDP>
static int attention_f32_hdp(const float* qkv, int cells, int tokens, int heads, int hd, float* out, cudaStream_t st) {
  const size_t smem = (size_t)2 * tokens * HDP * sizeof(float);
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_f32_kernel<HDP>), (int)(2 * 128 * HDP * 4), "cudaFuncSetAttribute(attention_f32_kernel)"));
  attention_f32_kernel<HDP><<<cells * heads, 128, smem, st>>>(qkv, tokens, heads, hd, out);
  RIBCA_LAUNCH_CHECK("attention_f32_kernel");
  return RIBCA_OK;
}

static int attention_f32(const float* qkv, int cells, int tokens, int heads, int hd, float* out, cudaStream_t st) {
  RIBCA_REQUIRE(tokens > 0 && tokens <= 128, "attention_f32: tokens=%d outside [1,128]", tokens);
  const int hdp = (hd + 15) / 16 * 16;
  switch (hdp) {
    case 16: return attention_f32_hdp<16>(qkv, cells, tokens, heads, hd, out, st);
    case 32: return attention_f32_hdp<32>(qkv, cells, tokens, heads, hd, out, st);
    case 48: return attention_f32_hdp<48>(qkv, cells, tokens, heads, hd, out, st);
    case 64: return attention_f32_hdp<64>(qkv, cells, tokens, heads, hd, out, st);
    default: set_error("attention_f32: unsupported head_dim %d", hd); return RIBCA_EUNSUPPORTED;
  }
}

// A[(cell*101 + 1 + py*10 + px)][c*16 + ky*4 + kx] = patch[cell][c][py*4+ky][px*4+kx]; class-token rows = 0
__global__ void __launch_bounds__(256)
im2col_f32_kernel(const float* __restrict__ patches, int n_cells, int C, float* __restrict__ a) {
  const int Kpe = 16 * C;
  const long long total = (long long)n_cells * C * 400;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int px = (int)(t % 10);
    const int y = (int)((t / 10) % 40);
    const int c = (int)((t / 400) % C);
    const long long cell = t / (400ll * C);
    const long long row = cell * 101 + 1 + (y >> 2) * 10 + px;
    *reinterpret_cast<float4*>(a + row * Kpe + c * 16 + (y & 3) * 4) = __ldg(reinterpret_cast<const float4*>(patches) + t);
  }
  const long long ztotal = (long long)n_cells * Kpe;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < ztotal; t += stride) {
    const long long cell = t / Kpe;
    a[cell * 101 * Kpe + (t - cell * Kpe)] = 0.0f;
  }
}

static int grid_rows(long long work, int per_block) {
  return (int)std::min<long long>((work + per_block - 1) / per_block, (long long)num_sms() * 16);
}

struct F32Buffers { float *x, *a, *qkv, *h; };

// timm Block x depth in fp32: x += proj(attn(LN1 x)); x += fc2(gelu(fc1(LN2 x)))
static int run_blocks_f32(const ribca_block_desc* blocks, int depth, int D, int heads, int cells, int tokens, const float* wf32,
                          const float* wmat, const F32Buffers& b, cudaStream_t st) {
  const int M = cells * tokens;
  const int hd = D / heads, hdp = (hd + 15) / 16 * 16;
  const int Wq = 3 * heads * hdp;
  for (int l = 0; l < depth; ++l) {
    const ribca_block_desc& w = blocks[l];
    layernorm_f32_kernel<<<grid_rows(M, 8), 256, 0, st>>>(b.x, M, D, wf32 + w.ln1_g, wf32 + w.ln1_b, 1e-6f, b.a);
    RIBCA_LAUNCH_CHECK("layernorm_f32_kernel");
    RIBCA_TRY(sgemm(b.a, wmat + w.qkv_w, M, Wq, D, wf32 + w.qkv_b, nullptr, 0, RIBCA_EPI_STORE, b.qkv, st));
    RIBCA_TRY(attention_f32(b.qkv, cells, tokens, heads, hd, b.a, st));
    RIBCA_TRY(sgemm(b.a, wmat + w.proj_w, M, D, D, wf32 + w.proj_b, nullptr, 0, RIBCA_EPI_RESIDUAL, b.x, st));
    layernorm_f32_kernel<<<grid_rows(M, 8), 256, 0, st>>>(b.x, M, D, wf32 + w.ln2_g, wf32 + w.ln2_b, 1e-6f, b.a);
    RIBCA_LAUNCH_CHECK("layernorm_f32_kernel");
    RIBCA_TRY(sgemm(b.a, wmat + w.fc1_w, M, 4 * D, D, wf32 + w.fc1_b, nullptr, 0, RIBCA_EPI_GELU, b.h, st));
    RIBCA_TRY(sgemm(b.h, wmat + w.fc2_w, M, D, 4 * D, wf32 + w.fc2_b, nullptr, 0, RIBCA_EPI_RESIDUAL, b.x, st));
  }
  return RIBCA_OK;
}

// the workspace is carved exactly like the tensor-core path's (stage4_networks.cu: x 4 B, a 2 x 2 B, qkv 4 B, h 2 x 2 B per
// element), so ribca_vit_workspace_bytes serves both
int vit_forward_f32(const ribca_vit_desc* desc, const float* wf32, const float* wmat, const float* patches, int n_cells,
                    float* probs, float* logits, float* x, float* a, float* qkv, float* h, cudaStream_t st) {
  const int D = desc->dim, T = desc->tokens, C = desc->in_chans;
  const int M = n_cells * T;
  const int Kpe = 16 * C;
  RIBCA_REQUIRE(Kpe <= 4 * D, "vit_forward_f32: patch-embed K=%d exceeds the MLP buffer", Kpe);
  im2col_f32_kernel<<<grid_rows((long long)n_cells * C * 400, 256), 256, 0, st>>>(patches, n_cells, C, h);
  RIBCA_LAUNCH_CHECK("im2col_f32_kernel");
  RIBCA_TRY(sgemm(h, wmat + desc->embed_w, M, D, Kpe, nullptr, wf32 + desc->embed_table, T, RIBCA_EPI_STORE, x, st));
  F32Buffers b{x, a, qkv, h};
  RIBCA_TRY(run_blocks_f32(desc->blocks, desc->depth, D, desc->heads, n_cells, T, wf32, wmat, b, st));
  return head_softmax_launch(x, n_cells, T, D, wf32 + desc->norm_g, wf32 + desc->norm_b, 1e-6f, wf32 + desc->head_w,
                             wf32 + desc->head_b, desc->classes, probs, logits, st);
}

}  // namespace ribca
