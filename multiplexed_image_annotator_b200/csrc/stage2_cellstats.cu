// Stage 2: label mask -> per-cell bbox / coordinate sums / area.
// Replaces ImageProcessor._cell_pos_dict (reference cta/preprocess.py:159-211) and the min / max /
// mean reductions its consumers apply to the pixel lists (cta/utils.py:227,232, cta/model.py:785-786).
// Integer arithmetic only: results are bit-exact whatever the accumulation order.
//
// HBM-bound: the int32 mask is read exactly once (4 B / pixel, 16-byte vector loads); each thread
// run-length-encodes 8 consecutive pixels of a row so a cell costs one group of atomics per
// (row, 8-pixel segment) instead of one per pixel.
#include "common.cuh"

namespace ribca {

constexpr int kPixPerThread = 8;

__global__ void mask_minmax_kernel(const int32_t* __restrict__ mask, long long n, int32_t* out2) {
  int lo = INT32_MAX, hi = INT32_MIN;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int v = __ldg(mask + i);
    lo = min(lo, v);
    hi = max(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out2, lo);
    atomicMax(out2 + 1, hi);
  }
}

__global__ void minmax_init_kernel(int32_t* out2) {
  out2[0] = INT32_MAX;
  out2[1] = INT32_MIN;
}

__global__ void stats_init_kernel(int32_t* bbox, unsigned long long* sums, int32_t* count, int n_ids) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ids) {
    reinterpret_cast<int4*>(bbox)[i] = make_int4(INT32_MAX, -1, INT32_MAX, -1);
    sums[2 * i] = 0ull;
    sums[2 * i + 1] = 0ull;
    count[i] = 0;
  }
}

__device__ __forceinline__ void flush_run(int id, int row, int c_first, int c_last, int max_id,
                                          int32_t* bbox, unsigned long long* sums, int32_t* count) {
  if (id <= 0 || id > max_id) return;
  int n = c_last - c_first + 1;
  atomicMin(&bbox[4 * id + 0], row);
  atomicMax(&bbox[4 * id + 1], row);
  atomicMin(&bbox[4 * id + 2], c_first);
  atomicMax(&bbox[4 * id + 3], c_last);
  atomicAdd(&sums[2 * id + 0], (unsigned long long)row * (unsigned long long)n);
  // sum of consecutive integers c_first..c_last
  atomicAdd(&sums[2 * id + 1], (unsigned long long)(c_first + c_last) * (unsigned long long)n / 2ull);
  atomicAdd(&count[id], n);
}

template <bool kVec>
__global__ void __launch_bounds__(256)
cell_stats_kernel(const int32_t* __restrict__ mask, int H, int W, int max_id, int32_t* bbox,
                  unsigned long long* sums, int32_t* count) {
  const int segs = (W + kPixPerThread - 1) / kPixPerThread;
  const long long total = (long long)H * segs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int row = (int)(t / segs);
    const int c0 = (int)(t % segs) * kPixPerThread;
    int v[kPixPerThread];
    const int32_t* p = mask + (long long)row * W + c0;
    if (kVec) {   // W % 8 == 0 and 16-byte aligned base: two 128-bit loads
      int4 a = __ldg(reinterpret_cast<const int4*>(p));
      int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < kPixPerThread; ++k) v[k] = (c0 + k < W) ? __ldg(p + k) : 0;
    }
    int run_id = v[0], run_start = c0;
#pragma unroll
    for (int k = 1; k < kPixPerThread; ++k) {
      if (v[k] != run_id) {
        flush_run(run_id, row, run_start, c0 + k - 1, max_id, bbox, sums, count);
        run_id = v[k];
        run_start = c0 + k;
      }
    }
    flush_run(run_id, row, run_start, c0 + kPixPerThread - 1, max_id, bbox, sums, count);
  }
}

// ---- compaction of the dense tables: labels with count > 0, ascending -------------------------
constexpr int kScanBlock = 1024;      // threads
constexpr int kScanItems = 4;         // ids per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__global__ void __launch_bounds__(kScanBlock)
compact_count_kernel(const int32_t* __restrict__ count, int n_ids, int* tile_sums) {
  __shared__ int total;
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int id = base + k;
    c += (id > 0 && id < n_ids && count[id] > 0) ? 1 : 0;
  }
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock)
compact_tilescan_kernel(int* tile_sums, int n_tiles, int32_t* n_cells) {
  // single block: exclusive scan of the tile totals, in place
  __shared__ int total;
  int carry = 0;
  for (int base = 0; base < n_tiles; base += kScanBlock) {
    int i = base + threadIdx.x;
    int v = (i < n_tiles) ? tile_sums[i] : 0;
    int ex = block_exclusive_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_cells = carry;
}

__global__ void __launch_bounds__(kScanBlock)
compact_write_kernel(const int32_t* __restrict__ bbox, const unsigned long long* __restrict__ sums,
                     const int32_t* __restrict__ count, int n_ids, const int* __restrict__ tile_sums,
                     int32_t* ids, int32_t* cbbox, unsigned long long* csums, int32_t* ccount,
                     int32_t* id_to_index) {
  __shared__ int total;
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int present[kScanItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int id = base + k;
    present[k] = (id > 0 && id < n_ids && count[id] > 0) ? 1 : 0;
    c += present[k];
  }
  int pos = tile_sums[blockIdx.x] + block_exclusive_scan(c, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int id = base + k;
    if (id < n_ids && id_to_index) id_to_index[id] = present[k] ? pos : -1;
    if (present[k]) {
      ids[pos] = id;
      reinterpret_cast<int4*>(cbbox)[pos] = reinterpret_cast<const int4*>(bbox)[id];
      csums[2 * pos] = sums[2 * id];
      csums[2 * pos + 1] = sums[2 * id + 1];
      ccount[pos] = count[id];
      ++pos;
    }
  }
}

// ---- CSR pixel lists ------------------------------------------------------------------------------
// cell_pos_dict of the reference (cta/preprocess.py:159-181, utils.py:272-290): per cell the row list and the column
// list of its pixels in raster order.  One warp per cell walks the cell's bbox window row by row, 32 columns at a time;
// a ballot + prefix popcount keeps the raster order.  rows / cols of cell j land at [offsets[j], offsets[j + 1]).
__global__ void __launch_bounds__(256)
cell_pixels_kernel(const int32_t* __restrict__ mask, int W, const int32_t* __restrict__ ids, const int32_t* __restrict__ cbbox,
                   const long long* __restrict__ offsets, int n_cells, int32_t* __restrict__ rows, int32_t* __restrict__ cols) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_cells; j += warps) {
    const int id = ids[j];
    const int4 bb = reinterpret_cast<const int4*>(cbbox)[j];          // rmin, rmax, cmin, cmax
    long long out = offsets[j];
    for (int r = bb.x; r <= bb.y; ++r) {
      const int32_t* row = mask + (long long)r * W;
      for (int c0 = bb.z; c0 <= bb.w; c0 += 32) {
        const int c = c0 + lane;
        const bool hit = c <= bb.w && row[c] == id;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const long long o = out + __popc(m & ((1u << lane) - 1u));
          rows[o] = r;
          cols[o] = c;
        }
        out += __popc(m);
      }
    }
  }
}

}  // namespace ribca

using namespace ribca;

extern "C" {

int ribca_cell_pixels(const int32_t* mask, int H, int W, const int32_t* ids, const int32_t* cbbox, const long long* offsets,
                      int n_cells, int32_t* rows, int32_t* cols, ribca_stream_t stream) {
  RIBCA_REQUIRE(mask && ids && cbbox && offsets && rows && cols, "ribca_cell_pixels: null pointer");
  RIBCA_REQUIRE(H > 0 && W > 0 && n_cells >= 0, "ribca_cell_pixels: bad shape");
  if (n_cells == 0) return RIBCA_OK;
  const int blocks = std::min((n_cells + 7) / 8, num_sms() * 16);
  cell_pixels_kernel<<<blocks, 256, 0, as_stream(stream)>>>(mask, W, ids, cbbox, offsets, n_cells, rows, cols);
  RIBCA_LAUNCH_CHECK("cell_pixels_kernel");
  return RIBCA_OK;
}

int ribca_mask_minmax(const int32_t* mask, long long n, int32_t* out2, ribca_stream_t stream) {
  RIBCA_REQUIRE(mask && out2 && n > 0, "ribca_mask_minmax: null pointer or empty mask");
  cudaStream_t st = as_stream(stream);
  minmax_init_kernel<<<1, 1, 0, st>>>(out2);
  RIBCA_LAUNCH_CHECK("minmax_init_kernel");
  int blocks = (int)std::min<long long>((n + 255) / 256, (long long)num_sms() * 16);
  mask_minmax_kernel<<<blocks, 256, 0, st>>>(mask, n, out2);
  RIBCA_LAUNCH_CHECK("mask_minmax_kernel");
  return RIBCA_OK;
}

int ribca_cell_stats(const int32_t* mask, int H, int W, int max_id, int32_t* bbox,
                     unsigned long long* sums, int32_t* count, ribca_stream_t stream) {
  RIBCA_REQUIRE(mask && bbox && sums && count, "ribca_cell_stats: null pointer");
  RIBCA_REQUIRE(H > 0 && W > 0 && max_id >= 0, "ribca_cell_stats: bad shape H=%d W=%d max_id=%d", H, W, max_id);
  cudaStream_t st = as_stream(stream);
  int n_ids = max_id + 1;
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_CELLSTATS, (double)H * (double)W * 4.0, st);
  stats_init_kernel<<<(n_ids + 255) / 256, 256, 0, st>>>(bbox, sums, count, n_ids);
  RIBCA_LAUNCH_CHECK("stats_init_kernel");
  long long total = (long long)H * ((W + kPixPerThread - 1) / kPixPerThread);
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
  bool vec = (W % kPixPerThread == 0) && ((reinterpret_cast<uintptr_t>(mask) & 15) == 0);
  if (vec)
    cell_stats_kernel<true><<<blocks, 256, 0, st>>>(mask, H, W, max_id, bbox, sums, count);
  else
    cell_stats_kernel<false><<<blocks, 256, 0, st>>>(mask, H, W, max_id, bbox, sums, count);
  if (prof) prof_end_span(st);
  RIBCA_LAUNCH_CHECK("cell_stats_kernel");
  return RIBCA_OK;
}

size_t ribca_compact_workspace_bytes(int max_id) {
  size_t tiles = ((size_t)max_id + 1 + kScanTile - 1) / kScanTile;
  return align_up(tiles * sizeof(int), 256);
}

int ribca_compact_cells(const int32_t* bbox, const unsigned long long* sums, const int32_t* count,
                        int max_id, int32_t* ids, int32_t* cbbox, unsigned long long* csums,
                        int32_t* ccount, int32_t* id_to_index, int32_t* n_cells, void* workspace,
                        size_t workspace_bytes, ribca_stream_t stream) {
  RIBCA_REQUIRE(bbox && sums && count && ids && cbbox && csums && ccount && n_cells && workspace,
                "ribca_compact_cells: null pointer");
  if (workspace_bytes < ribca_compact_workspace_bytes(max_id)) {
    set_error("ribca_compact_cells: workspace %zu < %zu", workspace_bytes, ribca_compact_workspace_bytes(max_id));
    return RIBCA_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  int n_ids = max_id + 1;
  int tiles = (n_ids + kScanTile - 1) / kScanTile;
  int* tile_sums = static_cast<int*>(workspace);
  compact_count_kernel<<<tiles, kScanBlock, 0, st>>>(count, n_ids, tile_sums);
  RIBCA_LAUNCH_CHECK("compact_count_kernel");
  compact_tilescan_kernel<<<1, kScanBlock, 0, st>>>(tile_sums, tiles, n_cells);
  RIBCA_LAUNCH_CHECK("compact_tilescan_kernel");
  compact_write_kernel<<<tiles, kScanBlock, 0, st>>>(bbox, sums, count, n_ids, tile_sums, ids, cbbox, csums,
                                                     ccount, id_to_index);
  RIBCA_LAUNCH_CHECK("compact_write_kernel");
  return RIBCA_OK;
}

}  // extern "C"
