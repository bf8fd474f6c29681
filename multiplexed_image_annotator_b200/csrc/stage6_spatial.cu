// Spatial statistics after the labels exist: exact k-nearest neighbours of the cell centroids and the
// two reductions the reference builds from them.
//   replaces  NearestNeighbors(n_neighbors=k, algorithm='ball_tree').fit(xy).kneighbors(xy)   cta/spatial_methods.py:35-40,97-101,153-155
//             the neighbourhood type matrix                                                   cta/spatial_methods.py:36-40,98-101
//             the multi-scale neighbour compositions of tissue_region_partition               cta/spatial_methods.py:157-176
// The reference queries a ball tree one point at a time from Python; here the centroids are binned into a uniform grid
// (host: one sort) and one thread per query walks the grid rings outwards from its bin, keeping the k best in a
// max-heap, until the k-th distance is inside the ring square already visited.  Distances are float64 with separate
// multiply / add roundings (no FMA contraction), i.e. the values sklearn compares; exact ties are broken by the lower
// point index (sklearn's tie order is unspecified).
#include "common.cuh"

namespace ribca {

constexpr int kKnnMaxK = 256;

struct KnnGrid {
  double x0, y0, inv_cell, cell;
  int gx, gy;
};

__device__ __forceinline__ bool knn_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

// xy_sorted: points in bin order; order[s] = original index of sorted point s; bin_start[b] .. bin_start[b+1]
__global__ void __launch_bounds__(128)
knn_grid_kernel(const double2* __restrict__ xy_sorted, const int* __restrict__ order, const int* __restrict__ bin_start, int n, int k,
                const KnnGrid g, int* __restrict__ out_idx, double* __restrict__ out_d2) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double hd[kKnnMaxK];
  int hi[kKnnMaxK];
  int cnt = 0;
  const double2 q = xy_sorted[s];
  const int bx = min(max((int)((q.x - g.x0) * g.inv_cell), 0), g.gx - 1);
  const int by = min(max((int)((q.y - g.y0) * g.inv_cell), 0), g.gy - 1);
  const int rmax = max(max(bx, g.gx - 1 - bx), max(by, g.gy - 1 - by));
  auto visit = [&](int xx, int yy) {
    if (xx < 0 || xx >= g.gx) return;
    const int b = yy * g.gx + xx;
    for (int p = bin_start[b]; p < bin_start[b + 1]; ++p) {
      const double2 c = xy_sorted[p];
      const double ex = c.x - q.x, ey = c.y - q.y;
      const double d2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
      const int id = order[p];
      if (cnt < k) {                                           // sift up
        int i = cnt++;
        while (i > 0) {
          const int par = (i - 1) >> 1;
          if (!knn_less(hd[par], hi[par], d2, id)) break;
          hd[i] = hd[par]; hi[i] = hi[par];
          i = par;
        }
        hd[i] = d2; hi[i] = id;
      } else if (knn_less(d2, id, hd[0], hi[0])) {             // replace the worst, sift down
        int i = 0;
        for (;;) {
          int ch = 2 * i + 1;
          if (ch >= k) break;
          if (ch + 1 < k && knn_less(hd[ch], hi[ch], hd[ch + 1], hi[ch + 1])) ++ch;
          if (!knn_less(d2, id, hd[ch], hi[ch])) break;
          hd[i] = hd[ch]; hi[i] = hi[ch];
          i = ch;
        }
        hd[i] = d2; hi[i] = id;
      }
    }
  };
  for (int r = 0; r <= rmax; ++r) {
    for (int dy = -r; dy <= r; ++dy) {                        // the bins at Chebyshev distance exactly r
      const int yy = by + dy;
      if (yy < 0 || yy >= g.gy) continue;
      if (dy == -r || dy == r) {
        for (int dx = -r; dx <= r; ++dx) visit(bx + dx, yy);
      } else {
        visit(bx - r, yy);
        visit(bx + r, yy);
      }
    }
    // every point outside the (2r+1)^2 square of bins is farther than r * cell from q
    const double reach = (double)r * g.cell;
    if (cnt == k && hd[0] <= reach * reach) break;
  }
  // heap sort: ascending (distance, index)
  for (int end = cnt - 1; end > 0; --end) {
    const double d = hd[end]; const int id = hi[end];
    hd[end] = hd[0]; hi[end] = hi[0];
    int i = 0;
    for (;;) {
      int ch = 2 * i + 1;
      if (ch >= end) break;
      if (ch + 1 < end && knn_less(hd[ch], hi[ch], hd[ch + 1], hi[ch + 1])) ++ch;
      if (!knn_less(d, id, hd[ch], hi[ch])) break;
      hd[i] = hd[ch]; hi[i] = hi[ch];
      i = ch;
    }
    hd[i] = d; hi[i] = id;
  }
  const long long o = (long long)order[s] * k;
  for (int j = 0; j < k; ++j) {
    out_idx[o + j] = j < cnt ? hi[j] : -1;
    if (out_d2) out_d2[o + j] = j < cnt ? hd[j] : 0.0;
  }
}

// type_matrix[type[j]][type[nbr[j][m]]] += 1 for m in [skip, k)   (spatial_methods.py:36-40: indices[1:])
__global__ void __launch_bounds__(256)
neighbor_matrix_kernel(const int* __restrict__ nbr, const int* __restrict__ types, int n, int k, int skip, int n_types,
                       unsigned long long* __restrict__ matrix) {
  __shared__ unsigned int local[RIBCA_MAX_TYPES * RIBCA_MAX_TYPES];
  for (int i = threadIdx.x; i < n_types * n_types; i += blockDim.x) local[i] = 0;
  __syncthreads();
  const long long total = (long long)n * (k - skip);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int j = (int)(t / (k - skip)), m = skip + (int)(t - (long long)j * (k - skip));
    const int nb = nbr[(long long)j * k + m];
    if (nb >= 0) atomicAdd(&local[types[j] * n_types + types[nb]], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_types * n_types; i += blockDim.x)
    if (local[i]) atomicAdd(&matrix[i], (unsigned long long)local[i]);
}

// comp[j][l * n_types + t] = #{m in [skip, skip + level[l]) : type[nbr[j][m]] == t} / level[l]   (spatial_methods.py:157-176)
struct LevelList { int n; int level[16]; };
__global__ void __launch_bounds__(128)
neighbor_composition_kernel(const int* __restrict__ nbr, const int* __restrict__ types, int n, int k, int skip, int n_types,
                            const __grid_constant__ LevelList lv, double* __restrict__ comp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int hist[RIBCA_MAX_TYPES];
  for (int t = 0; t < n_types; ++t) hist[t] = 0;
  const int* row = nbr + (long long)j * k + skip;
  double* out = comp + (long long)j * lv.n * n_types;
  int m = 0;
  for (int l = 0; l < lv.n; ++l) {
    for (; m < lv.level[l]; ++m) {
      const int nb = row[m];
      if (nb >= 0) ++hist[types[nb]];
    }
    int tot = 0;
    for (int t = 0; t < n_types; ++t) tot += hist[t];
    for (int t = 0; t < n_types; ++t) out[l * n_types + t] = (double)hist[t] / (double)tot;       // temp /= np.sum(temp)
  }
}

}  // namespace ribca

using namespace ribca;

extern "C" {

int ribca_knn_2d(const double* xy_sorted, const int* order, const int* bin_start, int n, int k, double x0, double y0,
                 double cell, int gx, int gy, int* out_idx, double* out_d2, ribca_stream_t stream) {
  RIBCA_REQUIRE(xy_sorted && order && bin_start && out_idx, "ribca_knn_2d: null pointer");
  RIBCA_REQUIRE(n > 0 && k > 0 && k <= kKnnMaxK && k <= n, "ribca_knn_2d: need 0 < k <= min(n, %d), got n=%d k=%d", kKnnMaxK, n, k);
  RIBCA_REQUIRE(cell > 0.0 && gx > 0 && gy > 0 && (long long)gx * gy < (1ll << 31), "ribca_knn_2d: bad grid");
  KnnGrid g;
  g.x0 = x0; g.y0 = y0; g.cell = cell; g.inv_cell = 1.0 / cell; g.gx = gx; g.gy = gy;
  knn_grid_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(reinterpret_cast<const double2*>(xy_sorted), order, bin_start, n, k, g,
                                                                 out_idx, out_d2);
  RIBCA_LAUNCH_CHECK("knn_grid_kernel");
  return RIBCA_OK;
}

int ribca_neighbor_stats(const int* nbr, const int* types, int n, int k, int skip, int n_types, unsigned long long* type_matrix,
                         const int* h_levels, int n_levels, double* compositions, ribca_stream_t stream) {
  RIBCA_REQUIRE(nbr && types, "ribca_neighbor_stats: null pointer");
  RIBCA_REQUIRE(n > 0 && k > skip && skip >= 0 && n_types > 0 && n_types <= RIBCA_MAX_TYPES, "ribca_neighbor_stats: bad sizes");
  cudaStream_t st = as_stream(stream);
  if (type_matrix) {
    const long long total = (long long)n * (k - skip);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
    neighbor_matrix_kernel<<<blocks, 256, 0, st>>>(nbr, types, n, k, skip, n_types, type_matrix);
    RIBCA_LAUNCH_CHECK("neighbor_matrix_kernel");
  }
  if (compositions) {
    RIBCA_REQUIRE(h_levels && n_levels > 0 && n_levels <= 16, "ribca_neighbor_stats: need 1..16 neighbour levels");
    LevelList lv;
    lv.n = n_levels;
    for (int l = 0; l < n_levels; ++l) {
      RIBCA_REQUIRE(h_levels[l] > 0 && h_levels[l] <= k - skip && (l == 0 || h_levels[l] > h_levels[l - 1]),
                    "ribca_neighbor_stats: levels must be ascending and <= k - skip");
      lv.level[l] = h_levels[l];
    }
    neighbor_composition_kernel<<<(n + 127) / 128, 128, 0, st>>>(nbr, types, n, k, skip, n_types, lv, compositions);
    RIBCA_LAUNCH_CHECK("neighbor_composition_kernel");
  }
  return RIBCA_OK;
}

}  // extern "C"
