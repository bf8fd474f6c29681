// Stage 4 primitive: C[M,N] = A[M,K] . W[N,K]^T (+ fused epilogue) on the 5th-gen tensor cores.
//
// The reference runs its ViT / MAE linears as fp32 PyTorch GEMMs (cta/model.py:397-406,
// cta/markerImputer.py:186-232).  A single 16-bit pass cannot hold the 1e-3 probability tolerance, so every
// operand is stored as TWO 16-bit planes (csrc/common.cuh) and the product is accumulated in fp32 TMEM:
//   f16f8  (default)  plane 0 fp16, plane 1 e4m3 pairs:  one kind::f8f6f4 pass over the pairs (both first-order
//                     correction terms at once, K doubled) + one kind::f16 pass: two passes' worth of tensor time
//   bf16x3            planes {hi, lo}:  lo.hi + hi.lo + hi.hi, three kind::f16 passes
//   bf16x1            hi plane only (throughput / debug)
// Because the split is a data layout, the kernel is a plain K-major GEMM: every K block stages the four tiles
// {A_0, A_1, W_0, W_1} ONCE (two TMA boxes of depth 2 over the plane axis) and issues all products from them.
//
// Kernel shape (sm_100a, cta_group::2):
//   persistent grid, one CTA per SM (CTA pairs, see below), 320 threads = 10 warps
//     warp 0     TMA producer: 3-D tensor maps (K, rows, plane), 128B swizzle, box 64 x rows x planes (whole
//                128-byte lines per row request: half the L2 tag lookups of 64-byte rows, -12.7 % GEMM time),
//                3-stage shared ring of 64 KB, mbarrier expect_tx; out-of-range K / rows are zero-filled by
//                the TMA unit, so K need not be a multiple of 32 nor M of 128
//     warp 1     allocates 512 TMEM columns, one lane (of the pair's leader) issues tcgen05.mma (M=256, N=BN<=256, K=16)
//                from shared-memory descriptors; tcgen05.commit releases ring slots / publishes
//                the accumulator
//     warps 2-9  epilogue (two warps per TMEM lane quadrant, alternating column chunks): tcgen05.ld ->
//                bias / row table / GELU / split -> swizzled shared staging -> TMA bulk tensor store
//                (full cache lines, no LSU serialisation); the residual add is a TMA reduce-add
//                (cp.reduce.async.bulk.tensor .add.f32), so x is never read into the SM.
//                TMEM is double-buffered (2 x 256 columns): the epilogue of tile t overlaps the main
//                loop of tile t+1.  In f16f8 mode the main loop is bound by L2 -> shared-memory delivery
//                (~6300 B/clk chip-wide; 64 FLOP per L2 byte at 256 x 256 pair tiles), see profiles/r01h_experiments.md
//   CTAs run as PAIRS (cluster of 2, tcgen05 cta_group::2): a pair owns a 256 x BN output tile, each CTA
//   stages its own 128 rows of A and HALF of the W tile, and the leader CTA issues M = 256 MMAs that read
//   both CTAs' shared memory.  Per SM this halves the B traffic (L2 -> smem and smem -> tensor core): the
//   single-CTA M = 128 form needs 96 B/cycle of operand reads plus 64 B/cycle of TMA writes against a
//   128 B/cycle shared-memory port, which capped the tensor pipe at ~80 %.  All barriers of the main
//   loop live in the leader: both CTAs' TMA loads complete_tx there; tcgen05.commit multicasts the
//   slot-free / accumulator-ready signals to both CTAs; the peer's epilogue releases the accumulator by a
//   remote mbarrier arrive.
//   tiles are ordered m-major so the CTAs of one wave share A rows through L2 and W stays L2-resident.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace ribca {

constexpr int BM = 128;
#ifndef RIBCA_BK
#define RIBCA_BK 64
#endif
constexpr int BK = RIBCA_BK;        // 16-bit elements per K block: 32 = one SWIZZLE_64B row (64 bytes), 64 = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
#ifndef RIBCA_STAGES
#define RIBCA_STAGES (RIBCA_BK == 64 ? 3 : 6)
#endif
constexpr int kStages = RIBCA_STAGES;
#ifndef RIBCA_EPI_WARPS
#define RIBCA_EPI_WARPS 8
#endif
constexpr int kEpiWarps = RIBCA_EPI_WARPS;   // kEpiWarps / 4 per TMEM lane quadrant, interleaved column chunks
constexpr int kEpiPerQuad = kEpiWarps / 4;
#ifndef RIBCA_STAGING_BUFS
#define RIBCA_STAGING_BUFS 1
#endif
constexpr int kStagingBufs = RIBCA_STAGING_BUFS;      // per-warp staging tiles (1: the ring takes 192 of the 227 KB)
constexpr int kMaxBN = 256;
constexpr int kGemmThreads = 32 * (2 + kEpiWarps);
constexpr int kATile = BM * BK * 2;             // one plane of A:  16 KB at BK = 64
constexpr int kABytes = 2 * kATile;             // both planes:    32 KB
constexpr int kBBytesMax = 2 * (kMaxBN / 2) * BK * 2; // this CTA's half of W, both planes: 32 KB
constexpr int kStageBytes = kABytes + kBBytesMax;
#ifndef RIBCA_MMA_ORDER
#define RIBCA_MMA_ORDER 0       // f16f8: 0 = e4m3 / fp16 instructions alternate per 16 elements, 1 = K block in e4m3, then in fp16
#endif
#ifndef RIBCA_FORCE_CW16
#define RIBCA_FORCE_CW16 0
#endif
#ifndef RIBCA_STAGING_BYTES
#define RIBCA_STAGING_BYTES 4096
#endif
constexpr int kStagingBytes = RIBCA_STAGING_BYTES;   // per epilogue warp: 32 rows x 128 B (fp32 x 32 cols, or bf16 hi + lo tiles)
constexpr int kStagingTile = kStagingBytes / kStagingBufs;   // 2 buffers: 16-column chunks (CW = 16), one store in flight per buffer
static_assert(kStagingTile == 4096 || (kStagingTile == 2048 && RIBCA_FORCE_CW16), "a 2 KB staging tile holds 16-column chunks only");
constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = 512;
#ifndef RIBCA_OPERAND_L2_PROMOTION
#define RIBCA_OPERAND_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_256B      // A/B-tested against 128B and NONE: no difference
#endif
// (register allocation is per 4 warps: 14 warps -> 16 x 32 x 128 registers; 18 warps would be capped at 96)
// RIBCA_GEMM_MAXNREG=112 (a 12-byte spill) leaves room for two 256-thread LayerNorm blocks beside a resident GEMM CTA: the
// co-residency experiment of profiles/r02_interleave.md; the default keeps the 154 registers of the unconstrained build
#ifdef RIBCA_GEMM_MAXNREG
#define RIBCA_GEMM_BOUNDS __maxnreg__(RIBCA_GEMM_MAXNREG)
#else
#define RIBCA_GEMM_BOUNDS __launch_bounds__(kGemmThreads, 1)
#endif

// shared-memory matrix descriptor of an operand tile whose rows are BK 16-bit elements
__device__ __forceinline__ uint64_t make_smem_desc_k(uint32_t smem_addr) {
  return BK == 64 ? make_smem_desc(smem_addr) : make_smem_desc_sw64(smem_addr);
}

struct GemmEpilogue {
  const float* bias;        // [N] or null
  const float* row_table;   // [table_period][N] or null
  int table_period;
  int mode;                 // ribca_epilogue
  int chunk;                // columns per staged TMA store: 32, or 16 when BN % 32 != 0
  float* out_f32;           // [M][N]          (SIMT path only; the tcgen05 path stores through tmap_out)
  __nv_bfloat16* out_hi;    // split output planes
  __nv_bfloat16* out_lo;
  int out_fmt;              // plane format of a split output (kFmtBf16 / kFmtF16F8)
  float acc_scale;          // accumulator scale (2^-(t+8) for f16f8 weights, else 1)
  // LayerNorm folded into this GEMM (include/ribca_b200.h: ribca_ln_fold)
  const float* stats_in;    // [M][kLnSlots][2] partial (sum, sum of squares) of the rows the A planes were split from, or null
  const float* c1;          // [N] sum_k W'[n][k] (stats_in != null); `bias` then holds c2
  int slots_in;             // filled slots of stats_in
  float inv_dim, ln_eps;    // 1 / D of the normalised rows, LayerNorm epsilon
  float* stats_out;         // *_LN epilogues: [M][kLnSlots][2] partial statistics of the stored fp32 rows
};
constexpr int kLnSlots = RIBCA_LN_SLOTS;

// mean and 1 / sqrt(var + eps) of one row from its partial sums (fixed slot order: the same bits on every schedule)
// the row's 8 slots are one 64-byte line: four independent 16-byte loads (one memory latency) ...
__device__ __forceinline__ void ln_stats_request(const float* __restrict__ stats, long long row, int M, float4 (&q)[kLnSlots / 2]) {
  const float4* p = reinterpret_cast<const float4*>(stats) + row * (kLnSlots / 2);
#pragma unroll
  for (int i = 0; i < kLnSlots / 2; ++i) q[i] = row < M ? __ldg(p + i) : make_float4(0.f, 0.f, 0.f, 0.f);
}
// ... then the fixed-order sum over the filled slots: mean and 1 / sqrt(var + eps) (the same bits on every schedule)
__device__ __forceinline__ void ln_stats_finish(const float4 (&q)[kLnSlots / 2], int slots, float inv_dim, float eps, float& mean, float& rstd) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnSlots / 2; ++i) {
    if (2 * i < slots) { s1 += q[i].x; s2 += q[i].y; }
    if (2 * i + 1 < slots) { s1 += q[i].z; s2 += q[i].w; }
  }
  mean = s1 * inv_dim;
  const float var = fmaxf(fmaf(-mean, mean, s2 * inv_dim), 0.f);
  rstd = 1.0f / sqrtf(var + eps);
}
__device__ __forceinline__ void ln_row_stats(const float* __restrict__ stats, long long row, int slots, float inv_dim, float eps,
                                             float& mean, float& rstd) {
  float4 q[kLnSlots / 2];
  ln_stats_request(stats, row, 0x7fffffff, q);
  ln_stats_finish(q, slots, inv_dim, eps, mean, rstd);
}

struct GemmShape {
  int M, N, K;
  int BN;                   // N tile (multiple of 16, <= 256, divides N)
  int n_planes;             // 2 (bf16x3 / f16f8: both planes staged) or 1 (bf16x1: hi only)
  int fmt;                  // operand plane format: kFmtBf16 or kFmtF16F8
};

// exact (erf) GELU of timm's Mlp.  erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7; measured GELU error
// < 5e-7 absolute in fp32, far below the split-bf16 output quantum) so the fc1 epilogue stays under the
// main loop: ~14 instructions instead of ~30 for erff.
__device__ __forceinline__ float gelu_erf(float x) {
  // a = |x| / sqrt(2) * sqrt(log2 e): exp(-(x/sqrt 2)^2) = 2^(-a^2); A-S's p is rescaled by 1 / sqrt(log2 e)
  const float a = fabsf(x) * 0.84932180028801904272f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2727374808792225f, a, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * a));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), erf_abs, hx);               // 0.5 x (1 + sign(x) erf|x/sqrt 2|)
}

// ---- epilogue: TMEM -> registers -> swizzled staging -> TMA store / reduce-add -----------------------
// CW columns per step.  fp32 outputs: staging rows of CW*4 bytes (128B swizzle for CW = 32, 64B for 16);
// split outputs: hi tile then lo tile, rows of CW*2 bytes (64B swizzle for CW = 32, 32B for 16).
template <int CW>
__device__ __forceinline__ void epilogue_loop(const CUtensorMap& tmap_out, const GemmShape& shp, const GemmEpilogue& epi,
                                              uint8_t* staging_base, uint32_t tmem_base, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, int warp, int lane) {
  const int BN = shp.BN;
  const int n_tiles_n = shp.N / BN;
  const int n_pairs = (((shp.M + BM - 1) / BM + 1) / 2) * n_tiles_n;
  const int cta_rank = (int)cluster_ctarank();
  const int quad = warp & 3;                             // TMEM lane quadrant this warp may read
  const int sub = (warp - 2) >> 2;                       // which of the warps of the quadrant
  uint8_t* stg_base = staging_base + (warp - 2) * kStagingBytes;
  int stg_sel = 0;
  const bool split_out = epi.mode == RIBCA_EPI_GELU || epi.mode == RIBCA_EPI_STORE_SPLIT;
  const int n_chunks = BN / CW;
  // folded LayerNorm: bv[0..3] = c1, bv[4..7] = c2 of 16 columns, requested half a chunk (or a tile) ahead
  float4 bv[8];
  bool c_ready = false;
  auto load_c = [&](int colx) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      bv[q] = __ldg(reinterpret_cast<const float4*>(epi.c1 + colx) + q);
      bv[4 + q] = __ldg(reinterpret_cast<const float4*>(epi.bias + colx) + q);
    }
  };
  float4 ln_q[kLnSlots / 2];
  if (epi.stats_in && (int)(blockIdx.x >> 1) < n_pairs)
    ln_stats_request(epi.stats_in, (2 * ((int)(blockIdx.x >> 1) / n_tiles_n) + cta_rank) * BM + quad * 32 + lane, shp.M, ln_q);
  int local = 0;
  for (int pr = blockIdx.x >> 1; pr < n_pairs; pr += gridDim.x >> 1, ++local) {
    const int buf = local & 1;
    const uint32_t use = (uint32_t)(local >> 1);
    const int m0 = (2 * (pr / n_tiles_n) + cta_rank) * BM, n0 = (pr % n_tiles_n) * BN;
    const int row = m0 + quad * 32 + lane;
    const float* table_row = epi.row_table ? epi.row_table + (long long)(row % epi.table_period) * shp.N : nullptr;
    // folded LayerNorm of the A rows: v = rstd * (acc * sc - mean * c1[col]) + c2[col]  ->  fma(acc, rs, fma(-mr, c1, c2))
    float ln_rs = 0.f, ln_mr = 0.f;
    if (epi.stats_in) {
      // this tile's row statistics were requested one tile ahead (a 64-byte line per row from a 26 MB table: a DRAM latency
      // that would otherwise sit in front of every tile); request the next tile's now
      float mean, rstd;
      ln_stats_finish(ln_q, epi.slots_in, epi.inv_dim, epi.ln_eps, mean, rstd);
      ln_rs = rstd * epi.acc_scale;
      ln_mr = mean * rstd;
      const int pn = pr + (gridDim.x >> 1);
      if (pn < n_pairs) ln_stats_request(epi.stats_in, (2 * (pn / n_tiles_n) + cta_rank) * BM + quad * 32 + lane, shp.M, ln_q);
    }
    mbar_wait(&tmem_full[buf], use & 1u);
    tcgen05_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kMaxBN);
    // the warps of a quadrant take interleaved column chunks; the starting warp rotates with the tile so that an
    // uneven chunk count (8 chunks over 3 warps) evens out
    int j0 = sub + local % kEpiPerQuad;
    if (j0 >= kEpiPerQuad) j0 -= kEpiPerQuad;
    for (int j = j0; j < n_chunks; j += kEpiPerQuad) {
      const int c = j * CW;
      const int col = n0 + c;
      uint8_t* stg = stg_base + stg_sel * kStagingTile;
      const uint32_t stg_addr = smem_u32(stg);
      if (kStagingBufs == 2) stg_sel ^= 1;
      float v[CW];
#pragma unroll
      for (int q = 0; q < CW / 16; ++q) tmem_ld16_nowait(t_row + (uint32_t)(c + 16 * q), reinterpret_cast<uint32_t*>(v) + 16 * q);
      // bias of this chunk: requested while the TMEM loads are in flight
      if (epi.stats_in) {
        if (!c_ready) load_c(col);
      } else if (epi.bias) {
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) bv[q] = __ldg(reinterpret_cast<const float4*>(epi.bias + col) + q);
      }
      tmem_ld_wait();
      const float sc = epi.acc_scale;                       // 1 except for f16f8 weights (2^-(t+8))
      // 16 columns at a time, each finished down to its shared-memory stores (short register live ranges)
#pragma unroll
      for (int hh = 0; hh < CW / 16; ++hh) {
        float* u = v + 16 * hh;
        if (epi.stats_in) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 a = bv[q], b = bv[4 + q];
            u[4 * q] = fmaf(u[4 * q], ln_rs, fmaf(-ln_mr, a.x, b.x)); u[4 * q + 1] = fmaf(u[4 * q + 1], ln_rs, fmaf(-ln_mr, a.y, b.y));
            u[4 * q + 2] = fmaf(u[4 * q + 2], ln_rs, fmaf(-ln_mr, a.z, b.z)); u[4 * q + 3] = fmaf(u[4 * q + 3], ln_rs, fmaf(-ln_mr, a.w, b.w));
          }
          // the vectors of the NEXT 16 columns (this chunk's second half, or the next chunk of the tile) are requested now: the
          // loads miss L1 (224 KB of it is shared memory) and used to sit on the long scoreboard in front of every half
          if (hh + 1 < CW / 16) load_c(col + 16 * (hh + 1));
          else if (j + kEpiPerQuad < n_chunks) { load_c(n0 + (j + kEpiPerQuad) * CW); c_ready = true; }
          else {
            // first chunk of this warp in the CTA's next tile (its start warp rotates with `local`)
            const int pn = pr + (gridDim.x >> 1);
            int jn = sub + (local + 1) % kEpiPerQuad;
            if (jn >= kEpiPerQuad) jn -= kEpiPerQuad;
            c_ready = pn < n_pairs && jn < n_chunks;
            if (c_ready) load_c((pn % n_tiles_n) * BN + jn * CW);
          }
        } else if (epi.bias) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = bv[4 * hh + q];
            u[4 * q] = fmaf(u[4 * q], sc, b.x); u[4 * q + 1] = fmaf(u[4 * q + 1], sc, b.y);
            u[4 * q + 2] = fmaf(u[4 * q + 2], sc, b.z); u[4 * q + 3] = fmaf(u[4 * q + 3], sc, b.w);
          }
        } else if (sc != 1.0f) {
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] *= sc;
        }
        if (table_row) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(table_row + col + 16 * hh) + q);
            u[4 * q] += b.x; u[4 * q + 1] += b.y; u[4 * q + 2] += b.z; u[4 * q + 3] += b.w;
          }
        }
        uint32_t hi[8], lo[8];
        if (split_out) {
          if (epi.mode == RIBCA_EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = gelu_erf(u[i]);
          }
          if (epi.out_fmt == kFmtF16F8) {
#pragma unroll
            for (int e = 0; e < 8; ++e) split_f16f8_x2(u[2 * e], u[2 * e + 1], hi[e], lo[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16x2(u[2 * e], u[2 * e + 1], hi[e], lo[e]);
          }
        }
        if (hh == 0) {
          // the math of the first half is done: only now make sure the store issued from this staging buffer
          // (kStagingBufs chunks ago) has been read
          if (lane == 0) {
            if (kStagingBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
        }
        if (split_out) {
          // rows of CW 16-bit elements (CW*2 bytes) per plane; 16-byte chunk index XOR-swizzled like the TMA store expects
          constexpr int kRowB = CW * 2;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const int ch = 2 * hh + c2;
            const int sw = (CW == 32) ? (ch ^ ((lane >> 1) & 3)) : (ch ^ ((lane >> 2) & 1));   // SWIZZLE_64B / SWIZZLE_32B
            const int off = lane * kRowB + (sw << 4);
            *reinterpret_cast<uint4*>(stg + off) = make_uint4(hi[4 * c2], hi[4 * c2 + 1], hi[4 * c2 + 2], hi[4 * c2 + 3]);
            *reinterpret_cast<uint4*>(stg + 32 * kRowB + off) = make_uint4(lo[4 * c2], lo[4 * c2 + 1], lo[4 * c2 + 2], lo[4 * c2 + 3]);
          }
        } else {
          constexpr int kRowB = CW * 4;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int ch = 4 * hh + c4;
            const int sw = (CW == 32) ? (ch ^ (lane & 7)) : (ch ^ ((lane >> 1) & 3));           // SWIZZLE_128B / SWIZZLE_64B
            *reinterpret_cast<float4*>(stg + lane * kRowB + (sw << 4)) = make_float4(u[4 * c4], u[4 * c4 + 1], u[4 * c4 + 2], u[4 * c4 + 3]);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int r0 = m0 + quad * 32;
        if (split_out) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(stg_addr), "r"(col), "r"(r0), "r"(0) : "memory");
        } else if (epi.mode == RIBCA_EPI_RESIDUAL) {
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(stg_addr), "r"(col), "r"(r0) : "memory");
        } else {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(stg_addr), "r"(col), "r"(r0) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(&tmem_empty[buf], 0);        // the leader's issuer owns the accumulator hand-off
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all stores of this warp have landed
  __syncwarp();
}


// ---- *_LN epilogues: the fp32 row store of a residual / embedding GEMM that ALSO leaves what the next LayerNorm-folded
// GEMM reads: the operand planes of the new rows and per-row partial sums (sum, sum of squares).  x_new = x_old + (acc * sc
// + bias) is computed in the SM, so memory holds exactly the value the planes and the statistics come from.  16-column
// chunks; the warp's 4 KB staging area is split in a 2 KB fp32 tile (32 rows x 64 B) and a 2 KB plane tile (hi, lo: 32 rows x
// 32 B each).  x_old / x_new move between global memory and the fp32 tile with COALESCED 16-byte accesses (lane l: row l / 4 +
// 8 i, 16-byte piece l % 4: 64-byte row segments, 8 rows per instruction) and between the tile and the row-per-lane
// accumulator layout through shared memory: row-per-lane global accesses (32 lines per instruction) were 73 % of the
// epilogue's stall samples (profiles/r02_lnfold.md).  x_old of the next chunk is requested one chunk ahead; the planes
// leave by TMA like every 16-bit output.  The two warps of a lane quadrant take the even / odd chunks in ascending order and
// own one statistics slot each: slot = (n0 / BN) * 2 + sub, so the partial sums do not depend on the schedule.
__device__ __forceinline__ void epilogue_loop_ln(const CUtensorMap& tmap_planes, const GemmShape& shp, const GemmEpilogue& epi,
                                                 uint8_t* staging_base, uint32_t tmem_base, uint64_t* tmem_full,
                                                 uint64_t* tmem_empty, int warp, int lane) {
  static_assert(kStagingBufs == 1 && kEpiPerQuad == 2 && kStagingBytes >= 4096, "one 4 KB staging area per warp; 2 warps per quadrant");
  constexpr int CW = 16;
  const int BN = shp.BN;
  const int n_tiles_n = shp.N / BN;
  const int n_pairs = (((shp.M + BM - 1) / BM + 1) / 2) * n_tiles_n;
  const int cta_rank = (int)cluster_ctarank();
  const int quad = warp & 3;
  const int sub = (warp - 2) >> 2;
  uint8_t* xs = staging_base + (warp - 2) * kStagingBytes;          // fp32 tile, rows of 64 B, 16-byte pieces XOR-swizzled
  uint8_t* ps = xs + 2048;                                          // plane tiles (TMA store source, SWIZZLE_32B)
  const uint32_t ps_addr = smem_u32(ps);
  const bool resid = epi.mode == RIBCA_EPI_RESIDUAL_LN;
  const int n_chunks = BN / CW;
  const float sc = epi.acc_scale;
  // row-per-lane view of the fp32 tile: lane = row, piece ch at ((ch ^ ((lane >> 1) & 3)) << 4)
  const int rp_base = lane * 64, rp_x = (lane >> 1) & 3;
  // coalesced view: piece l % 4 of rows l / 4 + 8 i
  const int co_piece = lane & 3, co_row = lane >> 2;
  int local = 0;
  for (int pr = blockIdx.x >> 1; pr < n_pairs; pr += gridDim.x >> 1, ++local) {
    const int buf = local & 1;
    const uint32_t use = (uint32_t)(local >> 1);
    const int m0 = (2 * (pr / n_tiles_n) + cta_rank) * BM, n0 = (pr % n_tiles_n) * BN;
    const int r0 = m0 + quad * 32;
    const int row = r0 + lane;
    const bool row_ok = row < shp.M;
    const float* table_row = epi.row_table ? epi.row_table + (long long)(row % epi.table_period) * shp.N : nullptr;
    float4 xg[4], bq[4];                                             // x_old (coalesced layout) and bias of the coming chunk
    auto load_x = [&](int col) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + co_row + 8 * i;
        xg[i] = (resid && r < shp.M) ? __ldcg(reinterpret_cast<const float4*>(epi.out_f32 + (long long)r * shp.N + col) + co_piece)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) bq[q] = epi.bias ? __ldg(reinterpret_cast<const float4*>(epi.bias + col) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (sub < n_chunks) load_x(n0 + sub * CW);           // x_old does not depend on the accumulator: in flight during the wait
    if (resid && sub == 0) {
      // x_old of this CTA's NEXT tile -> L2 (one BN * 4-byte row segment per lane): the register prefetch below then covers an
      // L2 latency, not a DRAM one
      const int pn = pr + (gridDim.x >> 1);
      if (pn < n_pairs) {
        const int rn = (2 * (pn / n_tiles_n) + cta_rank) * BM + quad * 32 + lane;
        if (rn < shp.M)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                       ::"l"(epi.out_f32 + (long long)rn * shp.N + (pn % n_tiles_n) * BN), "r"(BN * 4) : "memory");
      }
    }
    mbar_wait(&tmem_full[buf], use & 1u);
    tcgen05_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kMaxBN);
    float s1 = 0.f, s2 = 0.f;
    for (int j = sub; j < n_chunks; j += kEpiPerQuad) {
      const int c = j * CW;
      const int col = n0 + c;
      float v[CW];
      tmem_ld16_nowait(t_row + (uint32_t)c, reinterpret_cast<uint32_t*>(v));
      // x_old: coalesced registers -> fp32 tile -> row per lane
      __syncwarp();                                        // the previous chunk's reads of the tile are done
      if (resid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = co_row + 8 * i;
          *reinterpret_cast<float4*>(xs + r * 64 + ((co_piece ^ ((r >> 1) & 3)) << 4)) = xg[i];
        }
      }
      __syncwarp();
      float4 xr[4], bc[4];
      if (resid) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) xr[ch] = *reinterpret_cast<const float4*>(xs + rp_base + ((ch ^ rp_x) << 4));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) bc[q] = bq[q];
      if (j + kEpiPerQuad < n_chunks) load_x(n0 + (j + kEpiPerQuad) * CW);      // next chunk's x_old, one chunk ahead
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // the plain epilogue's operation order: acc * sc + bias, + table, then the residual add (x += v): same roundings
        float* u = v + 4 * q;
        if (epi.bias) {
          const float4 b = bc[q];
          u[0] = fmaf(u[0], sc, b.x); u[1] = fmaf(u[1], sc, b.y); u[2] = fmaf(u[2], sc, b.z); u[3] = fmaf(u[3], sc, b.w);
        } else {
          u[0] *= sc; u[1] *= sc; u[2] *= sc; u[3] *= sc;
        }
        if (table_row) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(table_row + col) + q);
          u[0] += t.x; u[1] += t.y; u[2] += t.z; u[3] += t.w;
        }
        if (resid) { u[0] += xr[q].x; u[1] += xr[q].y; u[2] += xr[q].z; u[3] += xr[q].w; }
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
      // x_new: row per lane -> fp32 tile -> coalesced stores (each lane re-reads only what the warp wrote: __syncwarp suffices)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<float4*>(xs + rp_base + ((ch ^ rp_x) << 4)) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = co_row + 8 * i;
        const float4 o = *reinterpret_cast<const float4*>(xs + r * 64 + ((co_piece ^ ((r >> 1) & 3)) << 4));
        if (r0 + r < shp.M) reinterpret_cast<float4*>(epi.out_f32 + (long long)(r0 + r) * shp.N + col)[co_piece] = o;
      }
      // ---- operand planes of the same values -> plane tiles -> TMA store -------------------------------------
      uint32_t hi[CW / 2], lo[CW / 2];
      if (epi.out_fmt == kFmtF16F8) {
#pragma unroll
        for (int e = 0; e < CW / 2; ++e) split_f16f8_x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
      } else {
#pragma unroll
        for (int e = 0; e < CW / 2; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous chunk's planes have been read
      __syncwarp();
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int off = lane * 32 + ((ch ^ ((lane >> 2) & 1)) << 4);                     // SWIZZLE_32B
        *reinterpret_cast<uint4*>(ps + off) = make_uint4(hi[4 * ch], hi[4 * ch + 1], hi[4 * ch + 2], hi[4 * ch + 3]);
        *reinterpret_cast<uint4*>(ps + 1024 + off) = make_uint4(lo[4 * ch], lo[4 * ch + 1], lo[4 * ch + 2], lo[4 * ch + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmap_planes)), "r"(ps_addr), "r"(col), "r"(r0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (row_ok && epi.stats_out && sub < n_chunks)
      reinterpret_cast<float2*>(epi.stats_out)[(long long)row * kLnSlots + (n0 / BN) * kEpiPerQuad + sub] = make_float2(s1, s2);
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(&tmem_empty[buf], 0);
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}

__global__ void __cluster_dims__(2, 1, 1) RIBCA_GEMM_BOUNDS
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out2,
                    const GemmShape shp, const GemmEpilogue epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging_base = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging_base + kEpiWarps * kStagingBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = shp.BN;
  const int n_tiles_n = shp.N / BN;
  const int n_tiles_m = (shp.M + BM - 1) / BM;
  const int n_pairs = ((n_tiles_m + 1) / 2) * n_tiles_n;   // a pair = M tiles (2p, 2p+1) x one N block, one per CTA of the cluster
  const uint32_t cta_rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
  const int n_kb = (shp.K + BK - 1) / BK;
  const int n_iter = n_kb;
  const int w_tile = (BN / 2) * BK * 2;                  // one plane of this CTA's half of the W tile
  const uint32_t stage_tx = (uint32_t)(2 * shp.n_planes * (kATile + w_tile));   // both CTAs' loads land on the leader's barrier
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_out);
    if (epi.mode == RIBCA_EPI_STORE_LN || epi.mode == RIBCA_EPI_RESIDUAL_LN) prefetch_tmap(&tmap_out2);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 2 * kEpiWarps); }   // both CTAs' epilogues
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers are initialised before any multicast can reach them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int pr = pair0; pr < n_pairs; pr += pair_stride) {
        const int m0 = (2 * (pr / n_tiles_n) + (int)cta_rank) * BM;
        const int n0 = (pr % n_tiles_n) * BN + (int)cta_rank * (BN / 2);     // this CTA's half of the W tile
        for (int kb = 0; kb < n_iter; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);        // released for both CTAs by the leader's multicast commit
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_3d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, m0, 0);    // box depth = n_planes: hi tile, then lo tile
          tma_load_3d_2sm(sb, &tmap_w, &full_bar[stage], kb * BK, n0, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      const uint32_t idesc = shp.fmt == kFmtF16F8 ? make_instr_desc_fmt0(2 * BM, BN) : make_instr_desc(2 * BM, BN);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      const int last_steps = (shp.K - (n_kb - 1) * BK + UMMA_K - 1) / UMMA_K;
      for (int pr = pair0; pr < n_pairs; pr += pair_stride, ++local) {
        const int buf = local & 1;
        const uint32_t use = (uint32_t)(local >> 1);
        mbar_wait(&tmem_empty[buf], (use & 1u) ^ 1u);      // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kMaxBN);
        for (int it = 0; it < n_iter; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
          // K steps of this K block: the last block of a K that is not a multiple of 64 (288, 144, 240, 112 ...) holds
          // TMA zero fill beyond K - no instruction is spent on it
#ifdef RIBCA_NO_KSKIP
          const int nk = BK / UMMA_K;       // A/B: spend the instructions on the zero fill
#else
          const int nk = it == n_iter - 1 ? last_steps : BK / UMMA_K;
#endif
          if (shp.fmt == kFmtF16F8) {
            // e4m3 pair planes (both correction terms, K doubled: 32 bytes = one K = 32 instruction per 16 elements)
            // then the fp16 planes; one accumulator, one instruction descriptor (format code 0 = E4M3 = F16)
#if RIBCA_MMA_ORDER == 1
            // A/B variant: the whole K block in e4m3, then the whole K block in fp16 (profiles/r02_gemm_order.md)
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              if (k < nk)
                umma_e4m3_2sm(d_tmem, make_smem_desc_k(a_addr + kATile + k * UMMA_K * 2), make_smem_desc_k(b_addr + w_tile + k * UMMA_K * 2),
                              idesc, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              if (k < nk)
                umma_bf16_2sm(d_tmem, make_smem_desc_k(a_addr + k * UMMA_K * 2), make_smem_desc_k(b_addr + k * UMMA_K * 2), idesc, 1u);
#else
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k >= nk) break;
              const uint64_t a_0 = make_smem_desc_k(a_addr + k * UMMA_K * 2);
              const uint64_t a_1 = make_smem_desc_k(a_addr + kATile + k * UMMA_K * 2);
              const uint64_t w_0 = make_smem_desc_k(b_addr + k * UMMA_K * 2);
              const uint64_t w_1 = make_smem_desc_k(b_addr + w_tile + k * UMMA_K * 2);
              umma_e4m3_2sm(d_tmem, a_1, w_1, idesc, (it > 0 || k > 0) ? 1u : 0u);
              umma_bf16_2sm(d_tmem, a_0, w_0, idesc, 1u);
            }
#endif
          } else if (shp.n_planes == 2) {
            // lo.hi + hi.lo + hi.hi from the four staged tiles
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k >= nk) break;
              const uint64_t a_hi = make_smem_desc_k(a_addr + k * UMMA_K * 2);
              const uint64_t a_lo = make_smem_desc_k(a_addr + kATile + k * UMMA_K * 2);
              const uint64_t w_hi = make_smem_desc_k(b_addr + k * UMMA_K * 2);
              const uint64_t w_lo = make_smem_desc_k(b_addr + w_tile + k * UMMA_K * 2);
              umma_bf16_2sm(d_tmem, a_lo, w_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
              umma_bf16_2sm(d_tmem, a_hi, w_lo, idesc, 1u);
              umma_bf16_2sm(d_tmem, a_hi, w_hi, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              if (k < nk)
                umma_bf16_2sm(d_tmem, make_smem_desc_k(a_addr + k * UMMA_K * 2), make_smem_desc_k(b_addr + k * UMMA_K * 2),
                              idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm_mcast(&empty_bar[stage], (uint16_t)0x3);   // slot reusable in both CTAs once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm_mcast(&tmem_full[buf], (uint16_t)0x3);   // accumulator complete: both CTAs' epilogues
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    if (epi.mode == RIBCA_EPI_STORE_LN || epi.mode == RIBCA_EPI_RESIDUAL_LN) {
      epilogue_loop_ln(tmap_out2, shp, epi, staging_base, tmem_base, tmem_full, tmem_empty, warp, lane);
    } else if (epi.chunk == 32) epilogue_loop<32>(tmap_out, shp, epi, staging_base, tmem_base, tmem_full, tmem_empty, warp, lane);
    else                        epilogue_loop<16>(tmap_out, shp, epi, staging_base, tmem_base, tmem_full, tmem_empty, warp, lane);
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                 // no CTA exits while its peer can still multicast into it
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Same contraction on the FP32 pipe (x = hi + lo reconstructed exactly, fp32 FMA accumulate):
// the on-device cross-check of the tensor-core kernel and the RIBCA_SIMT_FP32 precision mode.
// ---------------------------------------------------------------------------------------------
constexpr int ST = 64, SK = 16;
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, long long a_plane, const __nv_bfloat16* __restrict__ W,
                 long long w_plane, const GemmShape shp, const GemmEpilogue epi) {
  __shared__ float As[SK][ST + 1];
  __shared__ float Ws[SK][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < shp.K; k0 += SK) {
    for (int idx = threadIdx.x; idx < ST * SK; idx += 256) {
      const int r = idx / SK, k = idx % SK;
      float a = 0.f, w = 0.f;
      if (m0 + r < shp.M && k0 + k < shp.K) {
        const long long o = (long long)(m0 + r) * shp.K + k0 + k;
        a = __bfloat162float(A[o]) + __bfloat162float(A[a_plane + o]);
      }
      if (n0 + r < shp.N && k0 + k < shp.K) {
        const long long o = (long long)(n0 + r) * shp.K + k0 + k;
        w = __bfloat162float(W[o]) + __bfloat162float(W[w_plane + o]);
      }
      As[k][r] = a;
      Ws[k][r] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= shp.M) continue;
    float mean = 0.f, rstd = 1.f;
    if (epi.stats_in) ln_row_stats(epi.stats_in, row, epi.slots_in, epi.inv_dim, epi.ln_eps, mean, rstd);
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= shp.N) continue;
      float v = acc[i][j];
      if (epi.stats_in) v = fmaf(v, rstd, fmaf(-mean * rstd, epi.c1[col], epi.bias[col]));
      else if (epi.bias) v += epi.bias[col];
      if (epi.row_table) v += epi.row_table[(long long)(row % epi.table_period) * shp.N + col];
      const long long o = (long long)row * shp.N + col;
      if (epi.mode == RIBCA_EPI_GELU) {
        split_bf16(gelu_erf(v), epi.out_hi[o], epi.out_lo[o]);
      } else if (epi.mode == RIBCA_EPI_STORE_SPLIT) {
        split_bf16(v, epi.out_hi[o], epi.out_lo[o]);
      } else if (epi.mode == RIBCA_EPI_RESIDUAL || epi.mode == RIBCA_EPI_RESIDUAL_LN) {
        epi.out_f32[o] += v;        // (*_LN on this path: planes and statistics follow in ln_planes_stats_kernel)
      } else {
        epi.out_f32[o] = v;
      }
    }
  }
}

// SIMT companion of the *_LN epilogues (and their on-device cross-check): operand planes and the row statistics of fp32
// rows, one warp per row; the whole row goes into slot 0
__global__ void __launch_bounds__(256)
ln_planes_stats_kernel(const float* __restrict__ x, int M, int D, int fmt, __nv_bfloat16* __restrict__ p0, __nv_bfloat16* __restrict__ p1,
                       float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
    const float2* xr = reinterpret_cast<const float2*>(x + (long long)row * D);
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D / 2; i += 32) {
      const float2 v = xr[i];
      s1 += v.x + v.y;
      s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2));
      uint32_t a, b;
      split_pair(v.x, v.y, fmt, a, b);
      reinterpret_cast<uint32_t*>(p0 + (long long)row * D)[i] = a;
      reinterpret_cast<uint32_t*>(p1 + (long long)row * D)[i] = b;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) reinterpret_cast<float2*>(stats)[(long long)row * kLnSlots] = make_float2(s1, s2);
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) split_bf16(x[i], hi[i], lo[i]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// 3-D map over a two-plane operand: dims (K, rows, 2 planes), box (BK, box_rows, n_planes), rows of BK * 2 bytes swizzled
static int make_operand_map(CUtensorMap* map, const void* base, long long plane_elems, int rows, int K, int box_rows,
                            int n_planes) {
  auto encode = tensor_map_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return RIBCA_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)n_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, RIBCA_OPERAND_L2_PROMOTION,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d box_rows=%d plane=%lld", (int)r, rows, K, box_rows, plane_elems);
    return RIBCA_ECUDA;
  }
  return RIBCA_OK;
}

// tensor map of the output: fp32 [M][N] (2-D) or split-bf16 planes [2][M][N] (3-D, both planes per store)
static int make_output_map(CUtensorMap* map, bool split, void* base, long long plane_elems, int M, int N, int chunk) {
  auto encode = tensor_map_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return RIBCA_ECUDA; }
  CUresult r;
  cuuint32_t estr[3] = {1, 1, 1};
  if (split) {
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, 2};
    cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)plane_elems * 2};
    cuuint32_t box[3] = {(cuuint32_t)chunk, 32, 2};
    r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               chunk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {(cuuint32_t)chunk, 32};
    r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               chunk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(output) failed (%d) M=%d N=%d chunk=%d", (int)r, M, N, chunk); return RIBCA_ECUDA; }
  return RIBCA_OK;
}

int pick_bn(int N) {
  for (int bn = kMaxBN; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

// slots a *_LN epilogue of width N fills (tcgen05 path: two per N tile; SIMT path: one)
int gemm_ln_slots(int N, int precision) { return precision == RIBCA_SIMT_FP32 ? 1 : 2 * (N / pick_bn(N)); }

int gemm_launch(const void* A, long long a_plane, const void* W, long long w_plane, int M, int N, int K,
                const float* bias, const float* row_table, int table_period, int epilogue, float* out_f32,
                void* out_split, long long out_plane, int precision, int w_log2_scale, cudaStream_t stream,
                const ribca_ln_fold* ln) {
  RIBCA_REQUIRE(A && W, "gemm: null operand");
  const bool ln_out = epilogue == RIBCA_EPI_STORE_LN || epilogue == RIBCA_EPI_RESIDUAL_LN;
  const bool ln_in = ln && ln->stats_in;
  RIBCA_REQUIRE(!ln_out || (ln && ln->stats_out && out_f32 && out_split), "gemm: a *_LN epilogue needs out_f32, out_split and ln->stats_out");
  RIBCA_REQUIRE(!ln_in || (ln->c1 && bias && ln->slots_in > 0 && ln->slots_in <= RIBCA_LN_SLOTS && !row_table && !ln_out),
                "gemm: a LayerNorm-folded GEMM needs c1, c2 (as bias) and 1..%d statistics slots", RIBCA_LN_SLOTS);
  RIBCA_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  RIBCA_REQUIRE(N % 16 == 0 && K % 8 == 0, "gemm: N=%d must be a multiple of 16 and K=%d of 8", N, K);
  const bool split_out = epilogue == RIBCA_EPI_GELU || epilogue == RIBCA_EPI_STORE_SPLIT;
  RIBCA_REQUIRE(epilogue >= RIBCA_EPI_STORE && epilogue <= RIBCA_EPI_RESIDUAL_LN, "gemm: unknown epilogue %d", epilogue);
  RIBCA_REQUIRE(split_out ? (out_split != nullptr) : (out_f32 != nullptr), "gemm: output is null");
  RIBCA_REQUIRE(!row_table || table_period > 0, "gemm: row table needs a period");
  RIBCA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (a_plane * 2) % 16 == 0 && (w_plane * 2) % 16 == 0,
                "gemm: operands must be 16-byte aligned");
  GemmShape shp;
  memset(&shp, 0, sizeof(shp));
  shp.M = M; shp.N = N; shp.K = K;
  shp.BN = pick_bn(N);
  RIBCA_REQUIRE(shp.BN > 0, "gemm: no N tile for N=%d", N);
  shp.n_planes = precision == RIBCA_BF16X1 ? 1 : 2;
  shp.fmt = precision == RIBCA_F16F8 ? kFmtF16F8 : kFmtBf16;
  GemmEpilogue epi;
  epi.bias = bias; epi.row_table = row_table; epi.table_period = table_period > 0 ? table_period : 1;
  epi.mode = epilogue; epi.out_f32 = out_f32; epi.chunk = 32;
  epi.out_hi = static_cast<__nv_bfloat16*>(out_split);
  epi.out_lo = out_split ? static_cast<__nv_bfloat16*>(out_split) + out_plane : nullptr;
  // f16f8: a GELU output feeds the next GEMM (same format); a plain split store feeds attention (bf16 {hi, lo})
  epi.out_fmt = (precision == RIBCA_F16F8 && (epilogue == RIBCA_EPI_GELU || ln_out)) ? kFmtF16F8 : kFmtBf16;
  epi.stats_in = ln_in ? ln->stats_in : nullptr;
  epi.c1 = ln_in ? ln->c1 : nullptr;
  epi.slots_in = ln_in ? ln->slots_in : 0;
  epi.inv_dim = 1.0f / (float)K;
  epi.ln_eps = ln ? ln->eps : 0.f;
  epi.stats_out = ln_out ? ln->stats_out : nullptr;
  RIBCA_REQUIRE(w_log2_scale > -64 && w_log2_scale < 64, "gemm: weight scale exponent %d out of range", w_log2_scale);
  epi.acc_scale = precision == RIBCA_F16F8 ? ldexpf(1.0f, -w_log2_scale) : 1.0f;

  if (precision == RIBCA_SIMT_FP32) {   // bf16 {hi, lo} planes only
    dim3 grid((N + ST - 1) / ST, (M + ST - 1) / ST);
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(A), a_plane,
                                               static_cast<const __nv_bfloat16*>(W), w_plane, shp, epi);
    RIBCA_LAUNCH_CHECK("gemm_simt_kernel");
    if (ln_out) {
      ln_planes_stats_kernel<<<std::min((M + 7) / 8, num_sms() * 8), 256, 0, stream>>>(out_f32, M, N, kFmtBf16, epi.out_hi, epi.out_lo, ln->stats_out);
      RIBCA_LAUNCH_CHECK("ln_planes_stats_kernel");
    }
    return RIBCA_OK;
  }
  RIBCA_REQUIRE(precision == RIBCA_BF16X3 || precision == RIBCA_BF16X1 || precision == RIBCA_F16F8, "gemm: unknown precision %d", precision);
  CUtensorMap map_a, map_w, map_out, map_out2;
  memset(&map_out2, 0, sizeof(map_out2));
  RIBCA_TRY(make_operand_map(&map_a, A, a_plane, M, K, BM, shp.n_planes));
  RIBCA_TRY(make_operand_map(&map_w, W, w_plane, N, K, shp.BN / 2, shp.n_planes));      // each CTA of a pair stages half of the W tile
  epi.chunk = (shp.BN % 32 == 0 && !RIBCA_FORCE_CW16) ? 32 : 16;
  RIBCA_REQUIRE(!split_out || (out_plane * 2) % 16 == 0, "gemm: split output plane stride must be 16-byte aligned");
  RIBCA_TRY(make_output_map(&map_out, split_out, split_out ? out_split : (void*)out_f32, out_plane, M, N, epi.chunk));
  if (ln_out) {
    RIBCA_REQUIRE((out_plane * 2) % 16 == 0, "gemm: plane stride of the *_LN planes must be 16-byte aligned");
    RIBCA_TRY(make_output_map(&map_out2, true, out_split, out_plane, M, N, 16));      // the *_LN epilogue works in 16-column chunks
  }
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel), (int)(kSmemBytes), "cudaFuncSetAttribute(gemm_tcgen05_kernel)"));
  const int n_pairs = (((M + BM - 1) / BM + 1) / 2) * (N / shp.BN);
  const int grid = 2 * std::min(n_pairs, num_sms() / 2);
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_GEMM, 2.0 * (double)M * (double)N * (double)K, stream);
  gemm_tcgen05_kernel<<<grid, kGemmThreads, kSmemBytes, stream>>>(map_a, map_w, map_out, map_out2, shp, epi);
  if (prof) prof_end_span(stream);
  RIBCA_LAUNCH_CHECK("gemm_tcgen05_kernel");
  return RIBCA_OK;
}

// generic fp32 -> operand planes: A role (activations) or W role (weights scaled by 2^log2_scale, f16f8 only)
__global__ void split_planes_kernel(const float* __restrict__ x, long long n_pairs, int fmt, int w_role, float sc,
                                    uint32_t* __restrict__ p0, uint32_t* __restrict__ p1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
    const float2 v = reinterpret_cast<const float2*>(x)[i];
    uint32_t a, b;
    if (fmt == kFmtF16F8 && w_role) split_f16f8_w_x2(v.x, v.y, sc, a, b);
    else split_pair(v.x, v.y, fmt, a, b);
    p0[i] = a;
    p1[i] = b;
  }
}

int split_planes_launch(const float* x, long long n, int fmt, int w_role, int log2_scale, void* p0, void* p1, cudaStream_t stream) {
  if (n <= 0) return RIBCA_OK;
  RIBCA_REQUIRE(n % 2 == 0, "split_planes: element count %lld must be even", n);
  RIBCA_REQUIRE(fmt == kFmtBf16 || fmt == kFmtF16F8, "split_planes: unknown format %d", fmt);
  RIBCA_REQUIRE(log2_scale > -64 && log2_scale < 64, "split_planes: scale exponent %d out of range", log2_scale);
  const long long pairs = n / 2;
  int blocks = (int)std::min<long long>((pairs + 255) / 256, (long long)num_sms() * 8);
  split_planes_kernel<<<blocks, 256, 0, stream>>>(x, pairs, fmt, w_role, ldexpf(1.0f, log2_scale), static_cast<uint32_t*>(p0),
                                                  static_cast<uint32_t*>(p1));
  RIBCA_LAUNCH_CHECK("split_planes_kernel");
  return RIBCA_OK;
}

int split_launch(const float* x, long long n, void* hi, void* lo, cudaStream_t stream) {
  if (n <= 0) return RIBCA_OK;
  int blocks = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  split_bf16_kernel<<<blocks, 256, 0, stream>>>(x, n, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
  RIBCA_LAUNCH_CHECK("split_bf16_kernel");
  return RIBCA_OK;
}

}  // namespace ribca

using namespace ribca;

extern "C" {

int ribca_gemm_splitbf16(const void* A, long long a_plane, const void* W, long long w_plane, int M, int N, int K,
                         const float* bias, const float* row_table, int table_period, int epilogue,
                         float* out_f32, void* out_split, long long out_plane, int precision, int w_log2_scale,
                         ribca_stream_t stream) {
  return gemm_launch(A, a_plane, W, w_plane, M, N, K, bias, row_table, table_period, epilogue, out_f32, out_split,
                     out_plane, precision, w_log2_scale, as_stream(stream), nullptr);
}

int ribca_gemm_ln(const void* A, long long a_plane, const void* W, long long w_plane, int M, int N, int K,
                  const float* bias, const float* row_table, int table_period, int epilogue, float* out_f32,
                  void* out_split, long long out_plane, int precision, int w_log2_scale, const ribca_ln_fold* ln,
                  ribca_stream_t stream) {
  return gemm_launch(A, a_plane, W, w_plane, M, N, K, bias, row_table, table_period, epilogue, out_f32, out_split,
                     out_plane, precision, w_log2_scale, as_stream(stream), ln);
}

int ribca_gemm_ln_slots(int N, int precision) {
  RIBCA_REQUIRE(N > 0 && N % 16 == 0 && ribca::pick_bn(N) > 0, "ribca_gemm_ln_slots: bad N=%d", N);
  return gemm_ln_slots(N, precision);
}

int ribca_split_planes(const float* x, long long n, int format, int w_role, int log2_scale, void* plane0, void* plane1,
                       ribca_stream_t stream) {
  RIBCA_REQUIRE(x && plane0 && plane1, "ribca_split_planes: null pointer");
  return split_planes_launch(x, n, format, w_role, log2_scale, plane0, plane1, as_stream(stream));
}

int ribca_split_bf16(const float* x, long long n, void* hi, void* lo, ribca_stream_t stream) {
  RIBCA_REQUIRE(x && hi && lo, "ribca_split_bf16: null pointer");
  return split_launch(x, n, hi, lo, as_stream(stream));
}

}  // extern "C"
