// Stage 4: multi-head self-attention of the ViT blocks on the 5th-gen tensor cores.
// (timm Attention inside the reference's VisionTransformer, cta/model.py:31-64: softmax(q k^T / sqrt(hd)) v.)
//
// Input: the QKV GEMM's output in split-bf16 planes [2][M][3 * heads * hdp] (hdp = head_dim rounded up
// to 16, padding columns are exact zeros because the padded weight rows are zero).
// One persistent CTA (128 threads) per SM keeps two (cell, head) items in flight; per item:
//   1. TMA: Q (128 rows), K and V (TP rows) of the head, hi and lo planes, 128B-swizzled 64-column boxes
//   2. S = Q K^T  : tcgen05.mma M=128, N=TP, K=hdp, three split passes (lo.hi + hi.lo + hi.hi) into TMEM
//   3. softmax    : thread r owns row r: tcgen05.ld -> scale, mask columns >= tokens, max, exp, sum;
//                   P is written back to shared memory as split-bf16 in the K-major swizzled layout
//                   (over the dead Q / K tiles)
//   4. O = P V    : tcgen05.mma M=128, N=hdp, K=TP; V is consumed as an MN-major B operand straight
//                   from its TMA tile (token rows, head_dim contiguous); three split passes
//   5. O / rowsum -> split-bf16 [2][M][D] (the next GEMM's A operand)
// Rows / columns beyond `tokens` inside the 128 x TP tile hold the next cell's tokens (or TMA zero
// fill); they are masked in the softmax and never stored.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace ribca {

typedef __nv_bfloat16 bf16;

constexpr int kAttThreads = 128;
constexpr int kQBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kSlotBytes = 65536 + 2 * 128 * 128;                  // Q/K (later P) region + V tiles, sized for TP = 128
constexpr int kAttSmemBytes = 2 * kSlotBytes + 1024 + 128;         // two items in flight + align + barriers

struct AttnParams {
  int cells, tokens, heads, hd, hdp, D;
  float scale;
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

// Two (cell, head) items are in flight per CTA (slots 0 / 1), each with its own shared tiles, TMEM
// accumulators and mbarriers, so the TMA loads and the MMAs of one slot run under the softmax of the other:
//   S0, S1 issued | softmax 0 -> PV0 issued | softmax 1 -> PV1 issued | O0 out, reload slot 0 | O1 out, reload slot 1
template <int TP>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const AttnParams p, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo) {
  constexpr int kKVBytes = TP * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kSlotBytes);
  uint64_t* bar_qk = bars;          // [2] TMA  -> S MMA
  uint64_t* bar_v = bars + 2;       // [2] TMA  -> PV MMA
  uint64_t* bar_s = bars + 4;       // [2] S done  -> softmax
  uint64_t* bar_o = bars + 6;       // [2] PV done -> output, slot reload
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;

  const uint32_t idesc_s = make_instr_desc(128, TP, false);
  const uint32_t idesc_o = make_instr_desc(128, p.hdp, true);
  const int ksteps_s = p.hdp / 16;
  const int n_items = p.cells * p.heads;
  const int my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  auto slot_q = [&](int sl, int pl) { return smem + sl * kSlotBytes + pl * kQBytes; };
  auto slot_k = [&](int sl, int pl) { return smem + sl * kSlotBytes + 2 * kQBytes + pl * kKVBytes; };
  auto slot_p = [&](int sl, int pl) { return smem + sl * kSlotBytes + pl * 32768; };      // overlays Q / K
  auto slot_v = [&](int sl, int pl) { return smem + sl * kSlotBytes + 65536 + pl * kKVBytes; };
  auto item_of = [&](int k) { return blockIdx.x + k * (int)gridDim.x; };                  // k-th item of this CTA

  auto issue_loads = [&](int sl, int item) {            // thread 0 only
    const int cell = item / p.heads, head = item - cell * p.heads;
    const int row0 = cell * p.tokens;
    const int cq = head * p.hdp, ck = (p.heads + head) * p.hdp, cv = (2 * p.heads + head) * p.hdp;
    mbar_expect_tx(&bar_qk[sl], 2u * kQBytes + 2u * kKVBytes);
    mbar_expect_tx(&bar_v[sl], 2u * kKVBytes);
    for (int pl = 0; pl < 2; ++pl) {
      tma_load_3d(slot_q(sl, pl), &tmap_q, &bar_qk[sl], cq, row0, pl);
      tma_load_3d(slot_k(sl, pl), &tmap_kv, &bar_qk[sl], ck, row0, pl);
    }
    for (int pl = 0; pl < 2; ++pl) tma_load_3d(slot_v(sl, pl), &tmap_kv, &bar_v[sl], cv, row0, pl);
  };
  auto issue_s = [&](int sl) {                           // thread 0 only: S = Q K^T, lo.hi + hi.lo + hi.hi
    const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
    uint32_t acc = 0;
    for (int ps = 0; ps < 3; ++ps) {
      const uint32_t qa = smem_u32(slot_q(sl, pa[ps])), kb = smem_u32(slot_k(sl, pb[ps]));
      for (int ks = 0; ks < ksteps_s; ++ks) {
        umma_bf16(tmem_base + sl * 128, make_smem_desc(qa + ks * 32), make_smem_desc(kb + ks * 32), idesc_s, acc);
        acc = 1;
      }
    }
    umma_commit(&bar_s[sl]);
  };
  auto issue_pv = [&](int sl) {                          // thread 0 only: O = P V
    const int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
    uint32_t acc = 0;
    for (int ps = 0; ps < 3; ++ps) {
      const uint32_t pa_addr = smem_u32(slot_p(sl, pa[ps])), vb = smem_u32(slot_v(sl, pb[ps]));
#pragma unroll
      for (int kk = 0; kk < TP / 16; ++kk) {
        umma_bf16(tmem_base + 256 + sl * 64, make_smem_desc(pa_addr + (kk >> 2) * 16384 + (kk & 3) * 32),
                  make_smem_desc_mn(vb + kk * 2048), idesc_o, acc);
        acc = 1;
      }
    }
    umma_commit(&bar_o[sl]);
  };
  // softmax of row `tid` of slot sl; P (split bf16, K-major 128B swizzle) over the dead Q / K tiles; returns the row sum
  auto softmax_to_p = [&](int sl) -> float {
    float s[TP];
#pragma unroll
    for (int c = 0; c < TP / 16; ++c) tmem_ld16(tmem_base + sl * 128 + lane_addr + c * 16, s + c * 16);
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      s[j] = j < p.tokens ? s[j] * p.scale : -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      s[j] = expf(s[j] - mx);            // exp(-inf) = 0 for the masked columns
      sum += s[j];
    }
    uint8_t* p_hi = slot_p(sl, 0);
    uint8_t* p_lo = slot_p(sl, 1);
    const int r = tid;
#pragma unroll
    for (int ch = 0; ch < TP / 8; ++ch) {
      __align__(16) bf16 h[8], l[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_bf16(s[ch * 8 + e], h[e], l[e]);
      const int off = (ch >> 3) * 16384 + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(p_hi + off) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(p_lo + off) = *reinterpret_cast<const uint4*>(l);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    return sum;
  };
  auto store_o = [&](int sl, int item, float sum) {
    const int cell = item / p.heads, head = item - cell * p.heads;
    const float inv = 1.0f / sum;
    const bool row_ok = tid < p.tokens;
    const long long ob = ((long long)(cell * p.tokens + tid)) * p.D + head * p.hd;
    for (int c = 0; c < p.hdp / 16; ++c) {
      float o[16];
      tmem_ld16(tmem_base + 256 + sl * 64 + lane_addr + c * 16, o);
      if (row_ok) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int d = c * 16 + q4 * 4;
          if (d < p.hd) {
            bf16 h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16(o[q4 * 4 + e] * inv, h[e], l[e]);
            *reinterpret_cast<uint2*>(out_hi + ob + d) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(out_lo + ob + d) = *reinterpret_cast<const uint2*>(l);
          }
        }
      }
    }
    tcgen05_fence_before();
  };

  if (tid == 0) {
    if (my_items > 0) issue_loads(0, item_of(0));
    if (my_items > 1) issue_loads(1, item_of(1));
  }
  uint32_t ph = 0;                                       // every barrier completes once per pair
  for (int k = 0; k < my_items; k += 2, ph ^= 1u) {
    const bool two = k + 1 < my_items;
    if (tid == 0) {
      mbar_wait(&bar_qk[0], ph);
      tcgen05_fence_after();
      issue_s(0);
      if (two) {
        mbar_wait(&bar_qk[1], ph);
        issue_s(1);
      }
    }
    float sum0, sum1 = 1.f;
    mbar_wait(&bar_s[0], ph);
    tcgen05_fence_after();
    sum0 = softmax_to_p(0);
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bar_v[0], ph);
      tcgen05_fence_after();
      issue_pv(0);
    }
    if (two) {
      mbar_wait(&bar_s[1], ph);
      tcgen05_fence_after();
      sum1 = softmax_to_p(1);
      __syncthreads();
      if (tid == 0) {
        mbar_wait(&bar_v[1], ph);
        tcgen05_fence_after();
        issue_pv(1);
      }
    }
    mbar_wait(&bar_o[0], ph);
    tcgen05_fence_after();
    if (tid == 0 && k + 2 < my_items) issue_loads(0, item_of(k + 2));     // slot 0's tiles are dead now
    store_o(0, item_of(k), sum0);
    if (two) {
      mbar_wait(&bar_o[1], ph);
      tcgen05_fence_after();
      if (tid == 0 && k + 3 < my_items) issue_loads(1, item_of(k + 3));
      store_o(1, item_of(k + 1), sum1);
    }
    __syncthreads();                     // every thread has drained S / O of this pair before the next MMAs overwrite them
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_qkv_map(CUtensorMap* map, const void* base, long long plane_elems, long long rows, int width, int box_rows) {
  auto encode = tensor_map_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return RIBCA_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention: cuTensorMapEncodeTiled failed (%d)", (int)r); return RIBCA_ECUDA; }
  return RIBCA_OK;
}

template <int TP>
static int launch_tc(const CUtensorMap& mq, const CUtensorMap& mkv, const AttnParams& p, bf16* hi, bf16* lo, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    RIBCA_TRY(check_cuda(cudaFuncSetAttribute(attention_tc_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes),
                         "cudaFuncSetAttribute(attention_tc_kernel)"));
    attr_set = true;
  }
  const int grid = std::min(p.cells * p.heads, num_sms());
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_ATTENTION, 4.0 * (double)p.cells * p.heads * (double)p.tokens * p.tokens * p.hd, st);
  attention_tc_kernel<TP><<<grid, kAttThreads, kAttSmemBytes, st>>>(mq, mkv, p, hi, lo);
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("attention_tc_kernel");
  return RIBCA_OK;
}

// qkv_split: [2][M][3*heads*hdp] bf16, plane stride qkv_plane elements; out_split [2][M][heads*hd]
int attention_tc_launch(const void* qkv_split, long long qkv_plane, int cells, int tokens, int heads, int hd,
                        void* out_split, long long out_plane, cudaStream_t st) {
  RIBCA_REQUIRE(tokens > 0 && tokens <= 128, "attention_tc: tokens=%d outside [1,128]", tokens);
  RIBCA_REQUIRE(hd > 0 && hd <= 64 && hd % 4 == 0, "attention_tc: head_dim=%d unsupported", hd);
  if (cells <= 0) return RIBCA_OK;
  AttnParams p;
  p.cells = cells; p.tokens = tokens; p.heads = heads; p.hd = hd; p.hdp = (hd + 15) / 16 * 16; p.D = heads * hd;
  p.scale = 1.0f / sqrtf((float)hd);
  const int width = 3 * heads * p.hdp;
  const long long M = (long long)cells * tokens;
  const int TP = tokens <= 112 ? 112 : 128;
  CUtensorMap mq, mkv;
  RIBCA_TRY(make_qkv_map(&mq, qkv_split, qkv_plane, M, width, 128));
  RIBCA_TRY(make_qkv_map(&mkv, qkv_split, qkv_plane, M, width, TP));
  bf16* hi = static_cast<bf16*>(out_split);
  bf16* lo = hi + out_plane;
  return TP == 112 ? launch_tc<112>(mq, mkv, p, hi, lo, st) : launch_tc<128>(mq, mkv, p, hi, lo, st);
}

}  // namespace ribca

extern "C" int ribca_attention_tc(const void* qkv_split, long long qkv_plane, int cells, int tokens, int heads,
                                  int head_dim, void* out_split, long long out_plane, ribca_stream_t stream) {
  RIBCA_REQUIRE(qkv_split && out_split && heads > 0, "ribca_attention_tc: bad arguments");
  return ribca::attention_tc_launch(qkv_split, qkv_plane, cells, tokens, heads, head_dim, out_split, out_plane,
                                    ribca::as_stream(stream));
}
