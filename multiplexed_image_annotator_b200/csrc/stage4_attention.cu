// Stage 4: multi-head self-attention of the ViT blocks on the 5th-gen tensor cores.
// (timm Attention inside the reference's VisionTransformer, cta/model.py:31-64: softmax(q k^T / sqrt(hd)) v.)
//
// Input: the QKV GEMM's output in split-bf16 planes [2][M][3 * heads * hdp] (hdp = head_dim rounded up
// to 16, padding columns are exact zeros because the padded weight rows are zero).
// One persistent CTA per SM with two independent warpgroups; each warpgroup streams (cell, head) items
// through its own shared-memory slot, TMEM columns and mbarriers:
//   1. TMA: Q (128 rows), K and V (TP rows) of the head, hi and lo planes, 128B-swizzled 64-column boxes
//   2. S = Q K^T  : tcgen05.mma M=128, N=TP, K=hdp, three split passes (lo.hi + hi.lo + hi.hi) into TMEM;
//                   as soon as it retires the next item's Q / K are prefetched into the same tiles
//   3. softmax    : thread r owns row r: tcgen05.ld -> max, ex2(scale*log2e*(s - max)), sum; P is packed
//                   to split bf16x2 and written back INTO TMEM over S (tcgen05.st), so it never touches
//                   shared memory
//   4. O = P V    : tcgen05.mma with A = P from TMEM and V as an MN-major B operand straight from its TMA
//                   tile (token rows, head_dim contiguous); three split passes; V is then prefetched
//   5. O / rowsum -> split-bf16 [2][M][D] (the next GEMM's A operand)
// While one warpgroup waits on its MMAs the other runs its softmax.  Rows / columns beyond `tokens`
// inside the 128 x TP tile hold the next cell's tokens (or TMA zero fill); they are masked and never stored.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <stdlib.h>

namespace ribca {

typedef __nv_bfloat16 bf16;

constexpr int kAttThreads = 512;
// per item group: Q, K, V x {hi, lo} tiles of TP rows x 128 B, + the O staging tile of the TMA store.
// (the M = 128 MMA reads 128 Q rows: rows TP..127 fall into the K tile that follows - finite garbage
// that only reaches S rows which are never used)
__host__ __device__ constexpr int att_slot_bytes(int tp) { return 8 * tp * 128; }
__host__ __device__ constexpr int att_smem_bytes(int tp) { return 2 * att_slot_bytes(tp) + 1024 + 128; }

struct AttnParams {
  int cells, tokens, heads, hd, hdp, D;
  float scale_log2e;                          // log2(e) / sqrt(head_dim)
  int out_fmt;                                // plane format of O (the proj GEMM's A operand)
};

// MN-major (N contiguous) B operand in a 128B-swizzled tile whose rows are K indices:
// cute canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> SBO = 1024 B between 8-row K groups
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused while N <= 64
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem]: A rows are TMEM lanes, each 32-bit column holds two consecutive K elements
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Thread mapping.  The CTA holds two ITEM GROUPS of 256 threads; a group streams (cell, head) items through its own
// shared-memory slot, 256 TMEM columns and mbarriers.  Inside a group TWO threads share a query row: warp w handles
// TMEM lane quadrant w % 4 (the hardware rule for tcgen05.ld / st) and column half w / 4 - S columns [0, 64) or
// [64, TP), O columns [0, HDP / 2) or [HDP / 2, HDP).  Half the per-thread instruction stream of the one-thread-per-row
// form and twice the warps per scheduler: the kernel was bound by the latency of one warp's dependent softmax chain with
// 2 warps per scheduler (issue 0.34, profiles/r01h_summary.md).  The two halves of a row exchange their maximum and
// their sum through two spare TMEM columns of the group (tcgen05.st / ld), ordered by the group's named barrier.
template <int TP, int HDP, int FMT>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_out, const AttnParams p, bf16* __restrict__ out_hi,
                    bf16* __restrict__ out_lo) {
  static_assert(TP == 112 && HDP % 16 == 0 && HDP <= 64, "tile shape");
  constexpr int kKVBytes = TP * 128;
  constexpr int kQBytes = TP * 128;
  constexpr int kSlotBytes = att_slot_bytes(TP);
  constexpr int kHalf0 = 64;                  // S columns of half 0; half 1 owns [64, TP)
  constexpr int kOHalf = HDP / 2;             // O columns per half: 8, 16, 24 or 32
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kSlotBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid >> 8;                  // item group = slot
  const int gt = tid & 255;                  // thread within the group
  const int half = gt >> 7;                  // column half
  const int row = gt & 127;                  // query row owned by this thread (shared with the other half's thread)
  const bool leader = gt == 0;
  uint64_t* bar_qk = bars + 4 * grp;         // TMA  -> S MMA
  uint64_t* bar_v = bar_qk + 1;              // TMA  -> PV MMA
  uint64_t* bar_s = bar_qk + 2;              // S done  -> softmax, Q / K reload
  uint64_t* bar_o = bar_qk + 3;              // PV done -> output, V reload
  if (tid == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    prefetch_tmap(&tmap_out);
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t tmem_sp = tmem_base + grp * 256;           // S (fp32, TP cols); later P_hi at +0, P_lo at +64 (bf16x2)
  const uint32_t tmem_o = tmem_base + grp * 256 + 128;      // O (fp32, HDP cols)
  const uint32_t tmem_x = tmem_base + grp * 256 + 192;      // exchange columns: max of half 0 / 1, sum of half 0 / 1

  uint8_t* slot = smem + grp * kSlotBytes;
  uint8_t* q_s[2] = {slot, slot + kQBytes};
  uint8_t* k_s[2] = {slot + 2 * kQBytes, slot + 2 * kQBytes + kKVBytes};
  uint8_t* v_s[2] = {slot + 4 * kQBytes, slot + 4 * kQBytes + kKVBytes};
  uint8_t* o_s = slot + 6 * kQBytes;         // [2 planes][tokens rows][hd] bf16, dense, for the TMA store
  const bool tma_out = (p.hd % 8) == 0;      // rows of hd * 2 bytes must be multiples of 16 B for the bulk store

  const uint32_t idesc_s = make_instr_desc(128, TP, false);
  const uint32_t idesc_o = make_instr_desc(128, HDP, true);
  constexpr int ksteps_s = HDP / 16;
  const int n_items = p.cells * p.heads;
  const int first = blockIdx.x * 2 + grp, stride = 2 * gridDim.x;
  const int my_items = first < n_items ? (n_items - first + stride - 1) / stride : 0;

  auto load_qk = [&](int item) {
    const int cell = item / p.heads, head = item - cell * p.heads;
    const int row0 = cell * p.tokens;
    mbar_expect_tx(bar_qk, 2u * kQBytes + 2u * kKVBytes);
    for (int pl = 0; pl < 2; ++pl) {
      tma_load_3d(q_s[pl], &tmap_q, bar_qk, head * HDP, row0, pl);
      tma_load_3d(k_s[pl], &tmap_kv, bar_qk, (p.heads + head) * HDP, row0, pl);
    }
  };
  auto load_v = [&](int item) {
    const int cell = item / p.heads, head = item - cell * p.heads;
    mbar_expect_tx(bar_v, 2u * kKVBytes);
    for (int pl = 0; pl < 2; ++pl) tma_load_3d(v_s[pl], &tmap_kv, bar_v, (2 * p.heads + head) * HDP, cell * p.tokens, pl);
  };

  if (leader && my_items > 0) { load_qk(first); load_v(first); }
  uint32_t ph = 0;
  for (int k = 0; k < my_items; ++k, ph ^= 1u) {
    const int item = first + k * stride;
    const int cell = item / p.heads, head = item - cell * p.heads;
    // ---- S = Q K^T ------------------------------------------------------------------------------------
    if (leader) {
      mbar_wait(bar_qk, ph);
      tcgen05_fence_after();
      const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};       // lo.hi, hi.lo, hi.hi
      uint32_t acc = 0;
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        const uint32_t qa = smem_u32(q_s[pa[ps]]), kb = smem_u32(k_s[pb[ps]]);
#pragma unroll
        for (int ks = 0; ks < ksteps_s; ++ks) {
          umma_bf16(tmem_sp, make_smem_desc(qa + ks * 32), make_smem_desc(kb + ks * 32), idesc_s, acc);
          acc = 1;
        }
      }
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tcgen05_fence_after();
    if (leader && k + 1 < my_items) load_qk(item + stride);     // Q / K tiles are dead: prefetch the next item
    // ---- softmax of this thread's half row; P -> TMEM over S ---------------------------------------------
    // half 0: columns [0, 64) = 4 chunks of 16; half 1: columns [64, 112) = 3 chunks
    constexpr int kMaxChunks = kHalf0 / 16;
    const int n_chunks = half == 0 ? kHalf0 / 16 : (TP - kHalf0) / 16;
    const int col0 = half * kHalf0;
    float s[kMaxChunks * 16];
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c)
      if (c < n_chunks) tmem_ld16_nowait(tmem_sp + lane_addr + col0 + c * 16, reinterpret_cast<uint32_t*>(s) + c * 16);
    tmem_ld_wait();
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      if (c < n_chunks) {
        if (col0 + c * 16 + 16 > p.tokens) {                      // only a chunk that crosses `tokens` needs the mask
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col0 + c * 16 + j >= p.tokens) s[c * 16 + j] = -INFINITY;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, s[c * 16 + j]);
      }
    }
    // exchange the half-row maxima (every row has at least its column 0 valid, so the joint maximum is finite)
    tmem_st1(tmem_x + lane_addr + half, __float_as_uint(mx));
    tmem_st_wait();
    tcgen05_fence_before();
    group_sync(grp);                       // also: every S value of the row is in registers before P overwrites S
    tcgen05_fence_after();
    {
      uint32_t m0, m1;
      tmem_ld2(tmem_x + lane_addr, m0, m1);
      tmem_ld_wait();
      mx = fmaxf(__uint_as_float(m0), __uint_as_float(m1));
    }
    const float off = mx * p.scale_log2e;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      if (c < n_chunks) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          s[c * 16 + j] = fast_exp2(fmaf(s[c * 16 + j], p.scale_log2e, -off));       // exp((s - max) / sqrt(hd)); 0 for masked columns
          s[c * 16 + j + 1] = fast_exp2(fmaf(s[c * 16 + j + 1], p.scale_log2e, -off));
          sum0 += s[c * 16 + j];
          sum1 += s[c * 16 + j + 1];
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16x2(s[c * 16 + 2 * e], s[c * 16 + 2 * e + 1], hi[e], lo[e]);
        tmem_st8(tmem_sp + lane_addr + (col0 >> 1) + c * 8, hi);
        tmem_st8(tmem_sp + lane_addr + 64 + (col0 >> 1) + c * 8, lo);
      }
    }
    const float sum = sum0 + sum1;
    tmem_st1(tmem_x + lane_addr + 2 + half, __float_as_uint(sum));
    tmem_st_wait();
    tcgen05_fence_before();
    group_sync(grp);
    // ---- O = P V --------------------------------------------------------------------------------------
    if (leader) {
      tcgen05_fence_after();
      mbar_wait(bar_v, ph);
      const uint32_t pa[3] = {64, 0, 0};                      // P_lo, P_hi, P_hi
      const int pb[3] = {0, 1, 0};                            // V_hi, V_lo, V_hi
      uint32_t acc = 0;
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        const uint32_t vb = smem_u32(v_s[pb[ps]]);
#pragma unroll
        for (int kk = 0; kk < TP / 16; ++kk) {
          umma_bf16_ts(tmem_o, tmem_sp + pa[ps] + kk * 8, make_smem_desc_mn(vb + kk * 2048), idesc_o, acc);
          acc = 1;
        }
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tcgen05_fence_after();
    if (leader && k + 1 < my_items) load_v(item + stride);      // V tiles are dead
    // ---- normalise, split, store: this thread's O columns [half * kOHalf, +kOHalf) ----------------------------
    {
      float o[kOHalf];
      uint32_t x0, x1;
      const uint32_t t_o = tmem_o + lane_addr + half * kOHalf;
      if constexpr (kOHalf == 8) {
        tmem_ld8_nowait(t_o, reinterpret_cast<uint32_t*>(o));
      } else if constexpr (kOHalf == 16) {
        tmem_ld16_nowait(t_o, reinterpret_cast<uint32_t*>(o));
      } else if constexpr (kOHalf == 24) {
        tmem_ld16_nowait(t_o, reinterpret_cast<uint32_t*>(o));
        tmem_ld8_nowait(t_o + 16, reinterpret_cast<uint32_t*>(o) + 16);
      } else {
        tmem_ld16_nowait(t_o, reinterpret_cast<uint32_t*>(o));
        tmem_ld16_nowait(t_o + 16, reinterpret_cast<uint32_t*>(o) + 16);
      }
      tmem_ld2(tmem_x + lane_addr + 2, x0, x1);
      tmem_ld_wait();
      const float inv = 1.0f / (__uint_as_float(x0) + __uint_as_float(x1));     // fixed order: the same bits in both halves
      const bool row_ok = row < p.tokens;
      const int d0 = half * kOHalf;                                              // first head-dim column of this thread
      if (tma_out) {
        // the previous item's bulk store must have finished reading the staging tile
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        group_sync(grp);
        if (row_ok) {
          const int row_b = p.hd * 2;
          uint8_t* dh = o_s + row * row_b;
          uint8_t* dl = o_s + p.tokens * row_b + row * row_b;
#pragma unroll
          for (int ch = 0; ch < kOHalf / 8; ++ch) {
            if (d0 + ch * 8 < p.hd) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_pair(o[ch * 8 + 2 * e] * inv, o[ch * 8 + 2 * e + 1] * inv, FMT, h[e], l[e]);
              *reinterpret_cast<uint4*>(dh + (d0 + ch * 8) * 2) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(dl + (d0 + ch * 8) * 2) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        group_sync(grp);
        if (leader) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(o_s)), "r"(head * p.hd), "r"(cell * p.tokens), "r"(0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        const long long ob = ((long long)(cell * p.tokens + row)) * p.D + head * p.hd;
        if (row_ok) {
#pragma unroll
          for (int q4 = 0; q4 < kOHalf / 4; ++q4) {
            const int d = d0 + q4 * 4;
            if (d < p.hd) {
              uint32_t h[2], l[2];
              split_pair(o[q4 * 4] * inv, o[q4 * 4 + 1] * inv, FMT, h[0], l[0]);
              split_pair(o[q4 * 4 + 2] * inv, o[q4 * 4 + 3] * inv, FMT, h[1], l[1]);
              *reinterpret_cast<uint2*>(out_hi + ob + d) = make_uint2(h[0], h[1]);
              *reinterpret_cast<uint2*>(out_lo + ob + d) = make_uint2(l[0], l[1]);
            }
          }
        }
        tcgen05_fence_before();
        group_sync(grp);
      }
    }
    // (the barrier inside the output phase orders this group's TMEM reads before the next item's MMAs)
  }
  if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// head_dim 12 / 24 / 32 / 48 (every classifier; 48 = vit_l, the headline): THREE item groups per CTA.
// The kernel above is bound by the serial chain of one (cell, head) item per group (TMA -> S MMA -> softmax -> PV MMA ->
// output, ~5 us) with only two items in flight per SM: DRAM at 56 % and the tensor pipe at 18 % (profiles/r01h_summary.md).
// A third item needs shared memory: the 128-byte-wide Q / K / V boxes hold 96 (or 64) bytes of head each.  Here Q and K are
// staged without padding: 32 columns as one tile of 64-byte rows (SWIZZLE_64B: K steps 0, 1) and, for head_dim 48, 16 more
// as a tile of 32-byte rows (SWIZZLE_32B: K step 2) - 21 KB (14 KB) per operand instead of 28 KB; V keeps its 128-byte
// MN-major tile (its last columns belong to the next head and are never read: N = head_dim).  The O staging tile of the TMA
// store lives in the V tiles, which are dead once the PV MMAs retire; the next V load is issued after that store has read
// them (by the MMA-issuing thread, right after it has issued the next S).  Slot = 70 KB (56 KB), three slots = 210 KB;
// TMEM: 160 columns per group (S / P 112, O <= 48).  One thread per query row (128 threads per group, 384 per CTA).
// 4096 cells x 12 heads of vit_l: 0.773 -> 0.683 ms = 5.58 TB/s algorithmic, 85 % of the HBM copy peak (profiles/r02_attention.md).
constexpr int kAtt3Threads = 384;
constexpr int kA3TP = 112;
constexpr int kA3T64 = kA3TP * 64;                   // 32-column tile of one plane
constexpr int kA3T32 = kA3TP * 32;                   // 16-column tile
constexpr int kA3V = kA3TP * 128;                    // V tile of one plane
__host__ __device__ constexpr int a3_slot_bytes(int hdp) {
  return (hdp >= 32 ? 4 * kA3T64 : 0) + 2 * kA3V + ((hdp == 48 || hdp == 16) ? 4 * kA3T32 : 0);
}
__host__ __device__ constexpr int a3_smem_bytes(int hdp) { return 3 * a3_slot_bytes(hdp) + 1024 + 256; }
static_assert(a3_slot_bytes(48) % 1024 == 0 && a3_slot_bytes(32) % 1024 == 0 && a3_slot_bytes(16) % 1024 == 0 && kA3T64 % 512 == 0 && kA3T32 % 256 == 0,
              "tile alignment of the swizzle modes");

// K-major, 32-byte swizzle (16 bf16 per row): 8-row groups are 256 B apart, layout type 6 (SWIZZLE_32B)
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
__device__ __forceinline__ void group128_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

template <int HDP, int FMT>
__global__ void __launch_bounds__(kAtt3Threads, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_64, const __grid_constant__ CUtensorMap tmap_32,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_out, const AttnParams p,
                     bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
  static_assert(HDP == 16 || HDP == 32 || HDP == 48, "Q / K pieces: 32 columns (HDP >= 32) and / or 16 columns (HDP 16, 48)");
  constexpr int TP = kA3TP;
  constexpr bool kPiece64 = HDP >= 32;
  constexpr bool kPiece32 = HDP == 48 || HDP == 16;
  constexpr int kCol32 = kPiece64 ? 32 : 0;                  // first column of the 16-column piece inside the head
  constexpr int kSlot = a3_slot_bytes(HDP);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kSlot);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = tid >> 7;                  // item group = slot (4 warps: TMEM lane quadrants 0..3)
  const int row = tid & 127;                 // query row owned by this thread
  const bool leader = row == 0;
  uint64_t* bar_qk = bars + 4 * grp;         // TMA  -> S MMA
  uint64_t* bar_v = bar_qk + 1;              // TMA  -> PV MMA
  uint64_t* bar_s = bar_qk + 2;              // S done  -> softmax, Q / K reload
  uint64_t* bar_o = bar_qk + 3;              // PV done -> output
  if (tid == 0) {
    if (kPiece64) prefetch_tmap(&tmap_64);
    if (kPiece32) prefetch_tmap(&tmap_32);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_out);
    for (int i = 0; i < 12; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const uint32_t tmem_sp = tmem_base + grp * 160;          // S (fp32, 112 cols); later P_hi at +0, P_lo at +56 (bf16x2, 56 cols each)
  const uint32_t tmem_o = tmem_sp + 112;                   // O (fp32, HDP cols)

  uint8_t* slot = smem + grp * kSlot;
  constexpr int k64Bytes = kPiece64 ? 4 * kA3T64 : 0;
  uint8_t* q64[2] = {slot, slot + kA3T64};
  uint8_t* k64[2] = {slot + 2 * kA3T64, slot + 3 * kA3T64};
  uint8_t* v_s[2] = {slot + k64Bytes, slot + k64Bytes + kA3V};
  uint8_t* t32 = slot + k64Bytes + 2 * kA3V;
  uint8_t* q32[2] = {t32, t32 + kA3T32};
  uint8_t* k32[2] = {t32 + 2 * kA3T32, t32 + 3 * kA3T32};
  uint8_t* o_s = v_s[0];                     // [2 planes][tokens rows][hd] bf16, dense: the TMA store's source, in the dead V tiles
  const bool tma_out = (p.hd % 8) == 0;      // rows of hd * 2 bytes must be multiples of 16 B for the bulk store (not head_dim 12)

  const uint32_t idesc_s = make_instr_desc(128, TP, false);
  const uint32_t idesc_o = make_instr_desc(128, HDP, true);
  const int n_items = p.cells * p.heads;
  const int first = blockIdx.x * 3 + grp, stride = 3 * gridDim.x;
  const int my_items = first < n_items ? (n_items - first + stride - 1) / stride : 0;

  auto load_qk = [&](int item) {
    const int cell = item / p.heads, head = item - cell * p.heads;
    const int row0 = cell * p.tokens;
    mbar_expect_tx(bar_qk, (kPiece64 ? 4u * kA3T64 : 0u) + (kPiece32 ? 4u * kA3T32 : 0u));
    for (int pl = 0; pl < 2; ++pl) {
      if (kPiece64) {
        tma_load_3d(q64[pl], &tmap_64, bar_qk, head * HDP, row0, pl);
        tma_load_3d(k64[pl], &tmap_64, bar_qk, (p.heads + head) * HDP, row0, pl);
      }
      if (kPiece32) {
        tma_load_3d(q32[pl], &tmap_32, bar_qk, head * HDP + kCol32, row0, pl);
        tma_load_3d(k32[pl], &tmap_32, bar_qk, (p.heads + head) * HDP + kCol32, row0, pl);
      }
    }
  };
  auto load_v = [&](int item) {
    const int cell = item / p.heads, head = item - cell * p.heads;
    mbar_expect_tx(bar_v, 2u * kA3V);
    for (int pl = 0; pl < 2; ++pl) tma_load_3d(v_s[pl], &tmap_v, bar_v, (2 * p.heads + head) * HDP, cell * p.tokens, pl);
  };

  if (leader && my_items > 0) { load_qk(first); load_v(first); }
  uint32_t ph = 0;
  for (int k = 0; k < my_items; ++k, ph ^= 1u) {
    const int item = first + k * stride;
    const int cell = item / p.heads, head = item - cell * p.heads;
    // ---- S = Q K^T ------------------------------------------------------------------------------------
    if (leader) {
      mbar_wait(bar_qk, ph);
      tcgen05_fence_after();
      const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};       // lo.hi, hi.lo, hi.hi
      uint32_t acc = 0;
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        if (kPiece64) {
          const uint32_t qa = smem_u32(q64[pa[ps]]), kb = smem_u32(k64[pb[ps]]);
          umma_bf16(tmem_sp, make_smem_desc_sw64(qa), make_smem_desc_sw64(kb), idesc_s, acc);
          umma_bf16(tmem_sp, make_smem_desc_sw64(qa + 32), make_smem_desc_sw64(kb + 32), idesc_s, 1u);
          acc = 1;
        }
        if (kPiece32)
          umma_bf16(tmem_sp, make_smem_desc_sw32(smem_u32(q32[pa[ps]])), make_smem_desc_sw32(smem_u32(k32[pb[ps]])), idesc_s, acc);
        acc = 1;
      }
      umma_commit(bar_s);
      if (k > 0 && tma_out) {
        // the previous item's output store has read the V tiles (its staging): only now may this item's V land there
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        load_v(item);
      }
    }
    mbar_wait(bar_s, ph);
    tcgen05_fence_after();
    if (leader && k + 1 < my_items) load_qk(item + stride);     // Q / K tiles are dead: prefetch the next item
    // ---- softmax of this thread's row; P -> TMEM over S -----------------------------------------------
    float s[TP];
#pragma unroll
    for (int c = 0; c < TP / 16; ++c) tmem_ld16_nowait(tmem_sp + lane_addr + c * 16, reinterpret_cast<uint32_t*>(s) + c * 16);
    tmem_ld_wait();
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      if (j >= 96 && j >= p.tokens) s[j] = -INFINITY;         // only the tail chunk can be out of range (96 < tokens <= 112)
      mx = fmaxf(mx, s[j]);
    }
    const float off = mx * p.scale_log2e;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < TP; j += 2) {
      s[j] = fast_exp2(fmaf(s[j], p.scale_log2e, -off));       // exp((s - max) / sqrt(hd)); 0 for masked columns
      s[j + 1] = fast_exp2(fmaf(s[j + 1], p.scale_log2e, -off));
      sum0 += s[j];
      sum1 += s[j + 1];
    }
    // every S value of the row is in registers: P may overwrite S (this thread's own TMEM lane only)
#pragma unroll
    for (int c = 0; c < TP / 16; ++c) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_bf16x2(s[c * 16 + 2 * e], s[c * 16 + 2 * e + 1], hi[e], lo[e]);
      tmem_st8(tmem_sp + lane_addr + c * 8, hi);
      tmem_st8(tmem_sp + lane_addr + 56 + c * 8, lo);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    group128_sync(grp);
    // ---- O = P V --------------------------------------------------------------------------------------
    if (leader) {
      tcgen05_fence_after();
      mbar_wait(bar_v, ph);
      tcgen05_fence_after();
      const uint32_t pa[3] = {56, 0, 0};                      // P_lo, P_hi, P_hi
      const int pb[3] = {0, 1, 0};                            // V_hi, V_lo, V_hi
      uint32_t acc = 0;
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        const uint32_t vb = smem_u32(v_s[pb[ps]]);
#pragma unroll
        for (int kk = 0; kk < TP / 16; ++kk) {
          umma_bf16_ts(tmem_o, tmem_sp + pa[ps] + kk * 8, make_smem_desc_mn(vb + kk * 2048), idesc_o, acc);
          acc = 1;
        }
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tcgen05_fence_after();
    // ---- normalise, split, stage in the (dead) V tiles, one TMA store per item ------------------------------------
    {
      const float inv = 1.0f / (sum0 + sum1);
      float o[HDP];
#pragma unroll
      for (int c = 0; c < HDP / 16; ++c) tmem_ld16_nowait(tmem_o + lane_addr + c * 16, reinterpret_cast<uint32_t*>(o) + c * 16);
      tmem_ld_wait();
      if (tma_out) {
        if (row < p.tokens) {
          const int row_b = p.hd * 2;
          uint8_t* dh = o_s + row * row_b;
          uint8_t* dl = o_s + p.tokens * row_b + row * row_b;
#pragma unroll
          for (int ch = 0; ch < HDP / 8; ++ch) {
            if (ch * 8 < p.hd) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_pair(o[ch * 8 + 2 * e] * inv, o[ch * 8 + 2 * e + 1] * inv, FMT, h[e], l[e]);
              *reinterpret_cast<uint4*>(dh + ch * 16) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(dl + ch * 16) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        fence_proxy_async_smem();
        tcgen05_fence_before();
        group128_sync(grp);                    // also: every thread's TMEM reads of O / P are done before the next item's MMAs
        if (leader) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(o_s)), "r"(head * p.hd), "r"(cell * p.tokens), "r"(0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        // head_dim 12: 24-byte rows are no TMA box; 8-byte stores straight from the registers, V is free at once
        if (leader && k + 1 < my_items) load_v(item + stride);
        if (row < p.tokens) {
          const long long ob = ((long long)(cell * p.tokens + row)) * p.D + head * p.hd;
#pragma unroll
          for (int q4 = 0; q4 < HDP / 4; ++q4) {
            if (q4 * 4 < p.hd) {
              uint32_t h[2], l[2];
              split_pair(o[q4 * 4] * inv, o[q4 * 4 + 1] * inv, FMT, h[0], l[0]);
              split_pair(o[q4 * 4 + 2] * inv, o[q4 * 4 + 3] * inv, FMT, h[1], l[1]);
              *reinterpret_cast<uint2*>(out_hi + ob + q4 * 4) = make_uint2(h[0], h[1]);
              *reinterpret_cast<uint2*>(out_lo + ob + q4 * 4) = make_uint2(l[0], l[1]);
            }
          }
        }
        tcgen05_fence_before();
        group128_sync(grp);
      }
    }
  }
  if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HDP, int FMT>
static int launch_tc3(const CUtensorMap& m64, const CUtensorMap& m32, const CUtensorMap& mv, const CUtensorMap& mo, const AttnParams& p,
                      bf16* hi, bf16* lo, cudaStream_t st) {
  constexpr int kSmem = a3_smem_bytes(HDP);
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_tc3_kernel<HDP, FMT>), kSmem, "cudaFuncSetAttribute(attention_tc3_kernel)"));
  const int grid = std::min((p.cells * p.heads + 2) / 3, num_sms());
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_ATTENTION, 4.0 * (double)p.cells * p.heads * (double)p.tokens * p.tokens * p.hd, st);
  attention_tc3_kernel<HDP, FMT><<<grid, kAtt3Threads, kSmem, st>>>(m64, m32, mv, mo, p, hi, lo);
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("attention_tc3_kernel");
  return RIBCA_OK;
}

// box of `box_cols` columns (64 / 32 / 16: one 128- / 64- / 32-byte swizzle row) x box_rows rows of one plane
static int make_qkv_map(CUtensorMap* map, const void* base, long long plane_elems, long long rows, int width, int box_rows,
                        int box_cols = 64) {
  auto encode = tensor_map_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return RIBCA_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention: cuTensorMapEncodeTiled failed (%d)", (int)r); return RIBCA_ECUDA; }
  return RIBCA_OK;
}

template <int TP, int HDP, int FMT>
static int launch_tc(const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap& mo, const AttnParams& p, bf16* hi, bf16* lo,
                     cudaStream_t st) {
  constexpr int kAttSmemBytes = att_smem_bytes(TP);
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_tc_kernel<TP, HDP, FMT>), (int)(kAttSmemBytes), "cudaFuncSetAttribute(attention_tc_kernel)"));
  const int grid = std::min((p.cells * p.heads + 1) / 2, num_sms());
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_ATTENTION, 4.0 * (double)p.cells * p.heads * (double)p.tokens * p.tokens * p.hd, st);
  attention_tc_kernel<TP, HDP, FMT><<<grid, kAttThreads, kAttSmemBytes, st>>>(mq, mkv, mo, p, hi, lo);
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("attention_tc_kernel");
  return RIBCA_OK;
}

template <int TP, int FMT>
static int launch_tc_hdp(const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap& mo, const AttnParams& p, bf16* hi, bf16* lo,
                         cudaStream_t st) {
  switch (p.hdp) {
    case 16: return launch_tc<TP, 16, FMT>(mq, mkv, mo, p, hi, lo, st);
    case 32: return launch_tc<TP, 32, FMT>(mq, mkv, mo, p, hi, lo, st);
    case 48: return launch_tc<TP, 48, FMT>(mq, mkv, mo, p, hi, lo, st);
    case 64: return launch_tc<TP, 64, FMT>(mq, mkv, mo, p, hi, lo, st);
    default: set_error("attention_tc: unsupported padded head_dim %d", p.hdp); return RIBCA_EUNSUPPORTED;
  }
}

// RIBCA_ATTN3=0 keeps the two-group kernel for head_dim 12 / 24 / 32 / 48 (A/B)
static bool three_groups_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RIBCA_ATTN3"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

// qkv_split: [2][M][3*heads*hdp] bf16, plane stride qkv_plane elements; out_split [2][M][heads*hd]
int attention_tc_launch(const void* qkv_split, long long qkv_plane, int cells, int tokens, int heads, int hd,
                        void* out_split, long long out_plane, int out_fmt, cudaStream_t st) {
  RIBCA_REQUIRE(out_fmt == kFmtBf16 || out_fmt == kFmtF16F8, "attention_tc: unknown plane format %d", out_fmt);
  RIBCA_REQUIRE(tokens > 0 && tokens <= 112, "attention_tc: tokens=%d outside [1,112]", tokens);
  RIBCA_REQUIRE(hd > 0 && hd <= 64 && hd % 4 == 0, "attention_tc: head_dim=%d unsupported", hd);
  if (cells <= 0) return RIBCA_OK;
  AttnParams p;
  p.cells = cells; p.tokens = tokens; p.heads = heads; p.hd = hd; p.hdp = (hd + 15) / 16 * 16; p.D = heads * hd;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)hd);
  p.out_fmt = out_fmt;
  const int width = 3 * heads * p.hdp;
  const long long M = (long long)cells * tokens;
  constexpr int TP = 112;
  CUtensorMap mq, mkv;
  RIBCA_TRY(make_qkv_map(&mq, qkv_split, qkv_plane, M, width, TP));
  RIBCA_TRY(make_qkv_map(&mkv, qkv_split, qkv_plane, M, width, TP));
  CUtensorMap mo;
  memset(&mo, 0, sizeof(mo));
  if (hd % 8 == 0) {      // output map: [2][M][D] bf16, one box = (head_dim cols, tokens rows, 2 planes), dense rows
    auto encode = tensor_map_encode_fn();
    cuuint64_t dims[3] = {(cuuint64_t)p.D, (cuuint64_t)M, 2};
    cuuint64_t strides[2] = {(cuuint64_t)p.D * 2, (cuuint64_t)out_plane * 2};
    cuuint32_t box[3] = {(cuuint32_t)hd, (cuuint32_t)tokens, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out_split, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("attention: cuTensorMapEncodeTiled(out) failed (%d)", (int)r); return RIBCA_ECUDA; }
  }
  bf16* hi = static_cast<bf16*>(out_split);
  bf16* lo = hi + out_plane;
  if (p.hdp <= 48 && (hd % 8 == 0 || p.hdp == 16) && tokens > 96 && three_groups_enabled()) {
    // head_dim 12 / 24 / 32 / 48: three item groups per CTA, unpadded Q / K tiles (attention_tc3_kernel)
    CUtensorMap m64, m32;
    RIBCA_TRY(make_qkv_map(&m64, qkv_split, qkv_plane, M, width, TP, 32));
    RIBCA_TRY(make_qkv_map(&m32, qkv_split, qkv_plane, M, width, TP, 16));
    const bool f8 = out_fmt == kFmtF16F8;
    if (p.hdp == 48) return f8 ? launch_tc3<48, kFmtF16F8>(m64, m32, mkv, mo, p, hi, lo, st) : launch_tc3<48, kFmtBf16>(m64, m32, mkv, mo, p, hi, lo, st);
    if (p.hdp == 32) return f8 ? launch_tc3<32, kFmtF16F8>(m64, m32, mkv, mo, p, hi, lo, st) : launch_tc3<32, kFmtBf16>(m64, m32, mkv, mo, p, hi, lo, st);
    return f8 ? launch_tc3<16, kFmtF16F8>(m64, m32, mkv, mo, p, hi, lo, st) : launch_tc3<16, kFmtBf16>(m64, m32, mkv, mo, p, hi, lo, st);
  }
  return out_fmt == kFmtF16F8 ? launch_tc_hdp<TP, kFmtF16F8>(mq, mkv, mo, p, hi, lo, st)
                              : launch_tc_hdp<TP, kFmtBf16>(mq, mkv, mo, p, hi, lo, st);
}

}  // namespace ribca

extern "C" int ribca_attention_tc(const void* qkv_split, long long qkv_plane, int cells, int tokens, int heads,
                                  int head_dim, void* out_split, long long out_plane, int format,
                                  ribca_stream_t stream) {
  RIBCA_REQUIRE(qkv_split && out_split && heads > 0, "ribca_attention_tc: bad arguments");
  return ribca::attention_tc_launch(qkv_split, qkv_plane, cells, tokens, heads, head_dim, out_split, out_plane, format,
                                    ribca::as_stream(stream));
}
