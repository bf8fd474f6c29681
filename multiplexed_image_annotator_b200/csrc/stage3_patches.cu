// Stage 3: per-cell 40x40 patch gather with the soft cell mask, written straight into the model
// input batches.  Replaces crop_cell + smooth (reference cta/utils.py:226-270) and
// ImageProcessor._img2patches (cta/preprocess.py:76-151) for cell_size = 30.
//
// One CTA per cell.  The soft mask follows the reference's arithmetic operation by operation so
// the patches are bit-identical to the scipy/numpy path:
//   m   = (window == id)                         40 rows x 40 bits, one uint64 per row
//   d_j = dilation(m, disk(j)), j = 1..4         shift/OR on the row words; pixels outside the
//                                                window never contribute (reflect == ignore for a disk)
//   s   = f32(m) ; s += d1 ; s += d2 ; s += G1(d2) ; s += d3 ; s += G1(d3) ; s += G2(d3) ;
//         s += d4 ; s += G1(d4) ; s += G2(d4) ; s += G3(d4)          (float32 running sum)
//   G_sigma = scipy.ndimage.gaussian_filter(float64, mode='nearest', truncate=4): axis 0 then
//         axis 1, each  acc = x[c]*w0 ; for k = r..1: acc += (x[c-k] + x[c+k]) * w[k]  in float64
//         with separate multiply and add (no FMA contraction), s = f32(f64(s) + G)
//   s /= 11 ; s /= max(s + 1e-6)                                       (float32)
//   value[c] = f32( f64(f32(img[c]) - min[c]) * f64(s) + f64(min[c]) ), zero outside the image
// HBM traffic per cell: 1600 * (4*C_read + 4) bytes read, 1600 * 4 * sum(C_panel) written
// (128-byte coalesced rows of the NCHW batch).
#include "common.cuh"

namespace ribca {

constexpr int P = RIBCA_PATCH;          // 40
constexpr int PP = P * P;               // 1600
constexpr int kThreads = 128;
constexpr unsigned long long kRowMask = (1ull << P) - 1ull;

struct PatchParams {
  float* out[RIBCA_MAX_PANELS];
  int n_ch[RIBCA_MAX_PANELS];
  int src[RIBCA_MAX_PANELS][RIBCA_MAX_PANEL_CH];   // image channel, or -1 = fill with -1
  int n_panels;
  double g[3][RIBCA_GAUSS_STRIDE];                 // half kernels sigma 1,2,3
};

__device__ __forceinline__ unsigned long long spread_bits(unsigned long long b, int w) {
  unsigned long long r = b;
  for (int dx = 1; dx <= w; ++dx) r |= (b << dx) | (b >> dx);
  return r & kRowMask;
}

// half-width of disk(j) at vertical offset |dy|:  floor(sqrt(j*j - dy*dy))
__device__ __forceinline__ int disk_halfwidth(int j, int ady) {
  int rem = j * j - ady * ady;
  int w = 0;
  while ((w + 1) * (w + 1) <= rem) ++w;
  return w;
}

// one separable Gaussian of the binary image `bits` (40 row words), added into s (float32)
__device__ __forceinline__ void add_gaussian(const unsigned long long* bits, const double* __restrict__ w,
                                             int r, double* tmp, float* s) {
  // axis 0 (along rows), 'nearest' boundary
  for (int p = threadIdx.x; p < PP; p += kThreads) {
    const int y = p / P, x = p - y * P;
    double acc = __dmul_rn((double)((bits[y] >> x) & 1ull), w[0]);
    for (int k = r; k >= 1; --k) {
      const int ya = max(y - k, 0), yb = min(y + k, P - 1);
      const double pair = __dadd_rn((double)((bits[ya] >> x) & 1ull), (double)((bits[yb] >> x) & 1ull));
      acc = __dadd_rn(acc, __dmul_rn(pair, w[k]));
    }
    tmp[p] = acc;
  }
  __syncthreads();
  // axis 1 (along columns)
  for (int p = threadIdx.x; p < PP; p += kThreads) {
    const int y = p / P, x = p - y * P;
    const double* row = tmp + y * P;
    double acc = __dmul_rn(row[x], w[0]);
    for (int k = r; k >= 1; --k) {
      const double pair = __dadd_rn(row[max(x - k, 0)], row[min(x + k, P - 1)]);
      acc = __dadd_rn(acc, __dmul_rn(pair, w[k]));
    }
    s[p] = (float)__dadd_rn((double)s[p], acc);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
build_patches_kernel(const float* __restrict__ img, const int32_t* __restrict__ mask, int C_img, int H, int W,
                     const float* __restrict__ min_val, const int32_t* __restrict__ ids,
                     const int32_t* __restrict__ cbbox, int cell_begin, int n_cells,
                     const __grid_constant__ PatchParams prm, double* __restrict__ avg_int,
                     int32_t* __restrict__ windows) {
  __shared__ unsigned long long m_bits[P];
  __shared__ unsigned long long any_bits[P];
  __shared__ unsigned long long d_bits[4][P];
  __shared__ float s[PP];
  __shared__ double tmp[PP];
  __shared__ float red[kThreads / 32];
  __shared__ int win[4];
  __shared__ double gw[3][RIBCA_GAUSS_STRIDE];

  const int j = blockIdx.x;            // cell within this launch
  if (j >= n_cells) return;
  const int cell = cell_begin + j;
  const int tid = threadIdx.x;
  const int id = ids[cell];

  if (tid == 0) {
    // window of utils.py:227-235 (patch 40: xm - 20 is an integer, so int() of the float is exact)
    const int4 bb = reinterpret_cast<const int4*>(cbbox)[cell];   // rmin rmax cmin cmax
    const int rm = (bb.x + bb.y) >> 1;     // floor division of a non-negative sum
    const int r0 = max(rm - P / 2, 0);
    const int r1 = min(r0 + P, H);
    const int cm = (bb.z + bb.w) >> 1;
    const int c0 = max(cm - P / 2, 0);
    const int c1 = min(c0 + P, W);
    win[0] = r0; win[1] = r1; win[2] = c0; win[3] = c1;
    if (windows) reinterpret_cast<int4*>(windows)[j] = make_int4(r0, r1, c0, c1);
  }
  if (tid < P) { m_bits[tid] = 0ull; any_bits[tid] = 0ull; }
  if (tid < 3 * RIBCA_GAUSS_STRIDE) gw[tid / RIBCA_GAUSS_STRIDE][tid % RIBCA_GAUSS_STRIDE] = prm.g[tid / RIBCA_GAUSS_STRIDE][tid % RIBCA_GAUSS_STRIDE];
  __syncthreads();
  const int r0 = win[0], r1 = win[1], c0 = win[2], c1 = win[3];
  const int wh = r1 - r0, ww = c1 - c0;

  // ---- mask window -> bit rows ------------------------------------------------------------------
  for (int p = tid; p < PP; p += kThreads) {
    const int y = p / P, x = p - y * P;
    int lab = 0;
    if (y < wh && x < ww) lab = __ldg(mask + (long long)(r0 + y) * W + (c0 + x));
    if (lab == id) atomicOr(&m_bits[y], 1ull << x);
    if (lab > 0) atomicOr(&any_bits[y], 1ull << x);
  }
  __syncthreads();

  // ---- four disk dilations ----------------------------------------------------------------------
  for (int t = tid; t < 4 * P; t += kThreads) {
    const int dj = t / P + 1, y = t % P;
    unsigned long long acc = 0ull;
    for (int dy = -dj; dy <= dj; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= P) continue;
      acc |= spread_bits(m_bits[yy], disk_halfwidth(dj, dy < 0 ? -dy : dy));
    }
    d_bits[dj - 1][y] = acc;
  }
  __syncthreads();

  // ---- float32 running sum in the reference's order ---------------------------------------------
  for (int p = tid; p < PP; p += kThreads) {
    const int y = p / P, x = p - y * P;
    s[p] = (float)((m_bits[y] >> x) & 1ull) + (float)((d_bits[0][y] >> x) & 1ull) + (float)((d_bits[1][y] >> x) & 1ull);
  }
  __syncthreads();
  add_gaussian(d_bits[1], gw[0], 4, tmp, s);                       // G1(d2)
  for (int p = tid; p < PP; p += kThreads) { const int y = p / P, x = p - y * P; s[p] = __fadd_rn(s[p], (float)((d_bits[2][y] >> x) & 1ull)); }
  __syncthreads();
  add_gaussian(d_bits[2], gw[0], 4, tmp, s);                       // G1(d3)
  add_gaussian(d_bits[2], gw[1], 8, tmp, s);                       // G2(d3)
  for (int p = tid; p < PP; p += kThreads) { const int y = p / P, x = p - y * P; s[p] = __fadd_rn(s[p], (float)((d_bits[3][y] >> x) & 1ull)); }
  __syncthreads();
  add_gaussian(d_bits[3], gw[0], 4, tmp, s);                       // G1(d4)
  add_gaussian(d_bits[3], gw[1], 8, tmp, s);                       // G2(d4)
  add_gaussian(d_bits[3], gw[2], 12, tmp, s);                      // G3(d4)

  // ---- s /= 11 ; s /= max(s + 1e-6) -------------------------------------------------------------
  float mx = 0.0f;
  for (int p = tid; p < PP; p += kThreads) {
    const float v = __fdiv_rn(s[p], 11.0f);
    s[p] = v;
    mx = fmaxf(mx, __fadd_rn(v, 1e-6f));
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k) mx = fmaxf(mx, red[k]);
  for (int p = tid; p < PP; p += kThreads) s[p] = __fdiv_rn(s[p], mx);
  __syncthreads();

  // ---- model inputs -----------------------------------------------------------------------------
  const long long plane = (long long)H * W;
  for (int pnl = 0; pnl < prm.n_panels; ++pnl) {
    const int nch = prm.n_ch[pnl];
    float* dst = prm.out[pnl] + (long long)j * nch * PP;
    for (int k = 0; k < nch; ++k) {
      const int src = prm.src[pnl][k];
      float* o = dst + (long long)k * PP;
      if (src < 0) {
        for (int p = tid; p < PP; p += kThreads) o[p] = -1.0f;
        continue;
      }
      const float mn = __ldg(min_val + src);
      const float* ip = img + src * plane;
      for (int p = tid; p < PP; p += kThreads) {
        const int y = p / P, x = p - y * P;
        float z = 0.0f;
        if (y < wh && x < ww) z = __fsub_rn(__ldg(ip + (long long)(r0 + y) * W + (c0 + x)), mn);
        o[p] = (float)__dadd_rn(__dmul_rn((double)z, (double)s[p]), (double)mn);
      }
    }
  }

  // ---- per-channel mean over every labelled pixel of the window (float64) -------------------------
  if (avg_int) {
    int n_sel = 0;
    for (int y = 0; y < P; ++y) n_sel += __popcll(any_bits[y]);
    const int lane = tid & 31, warp = tid >> 5;
    for (int c = warp; c < C_img; c += kThreads / 32) {
      const float mn = __ldg(min_val + c);
      const float* ip = img + c * plane;
      double acc = 0.0;
      for (int p = lane; p < PP; p += 32) {
        const int y = p / P, x = p - y * P;
        if (!((any_bits[y] >> x) & 1ull)) continue;
        float z = 0.0f;
        if (y < wh && x < ww) z = __fsub_rn(__ldg(ip + (long long)(r0 + y) * W + (c0 + x)), mn);
        acc += __dadd_rn(__dmul_rn((double)z, (double)s[p]), (double)mn);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) avg_int[(long long)j * C_img + c] = acc / (double)n_sel;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cell_size != 30: window edge PW = int(40 * cell_size / 30) (reference cta/preprocess.py:67,78), soft
// mask and marker values on the PW x PW window exactly as above, then skimage.transform.resize(order=0,
// anti_aliasing=True, preserve_range=True) of the (C_img, PW, PW) float64 patch to 40 x 40
// (cta/preprocess.py:106): when down-scaling a float64 Gaussian of sigma (PW/40 - 1)/2 along both image
// axes (scipy 'mirror' boundary, same tap order as above), nearest sampling at the source indices scipy's
// zoom(grid_mode=True) uses (computed on the host in float64 and passed in), and a clip to the
// [min, max] of the whole patch over ALL image channels.  Generic, byte-mask implementation.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPW = 80;
constexpr int kThreadsR = 256;

struct ResizeParams {
  int pw;                     // window edge
  int src[P];                 // source index of each of the 40 output rows / columns
  int r_aa;                   // anti-alias radius, -1 = none
  double w_aa[8];             // half kernel of the anti-alias Gaussian
};

__device__ __forceinline__ int mirror_index(int i, int n) {       // scipy 'mirror': d c b | a b c d | c b a
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  int m = i % p;
  if (m < 0) m += p;
  return m < n ? m : p - m;
}

__device__ __forceinline__ void add_gaussian_bytes(const unsigned char* d, int pw, const double* __restrict__ w, int r,
                                                   double* tmp, float* s) {
  const int n = pw * pw;
  for (int p = threadIdx.x; p < n; p += kThreadsR) {
    const int y = p / pw, x = p - y * pw;
    double acc = __dmul_rn((double)d[p], w[0]);
    for (int k = r; k >= 1; --k) {
      const int ya = max(y - k, 0), yb = min(y + k, pw - 1);
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)d[ya * pw + x], (double)d[yb * pw + x]), w[k]));
    }
    tmp[p] = acc;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += kThreadsR) {
    const int y = p / pw, x = p - y * pw;
    const double* row = tmp + y * pw;
    double acc = __dmul_rn(row[x], w[0]);
    for (int k = r; k >= 1; --k)
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(row[max(x - k, 0)], row[min(x + k, pw - 1)]), w[k]));
    s[p] = (float)__dadd_rn((double)s[p], acc);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreadsR)
build_patches_resized_kernel(const float* __restrict__ img, const int32_t* __restrict__ mask, int C_img, int H, int W,
                             const float* __restrict__ min_val, const int32_t* __restrict__ ids,
                             const int32_t* __restrict__ cbbox, int cell_begin, int n_cells,
                             const __grid_constant__ PatchParams prm, const __grid_constant__ ResizeParams rz,
                             double* __restrict__ avg_int, int32_t* __restrict__ windows) {
  extern __shared__ __align__(16) unsigned char dyn[];
  const int pw = rz.pw, n = pw * pw;
  double* tmp = reinterpret_cast<double*>(dyn);                   // [n]
  double* val = tmp + n;                                          // [n]
  float* s = reinterpret_cast<float*>(val + n);                   // [n]
  unsigned char* mk = reinterpret_cast<unsigned char*>(s + n);    // [n] bit0: label == id, bit1: label > 0
  unsigned char* dcur = mk + n;                                   // [n] current dilation
  __shared__ double red_lo[kThreadsR / 32], red_hi[kThreadsR / 32], red_sum[kThreadsR / 32];
  __shared__ float red_f[kThreadsR / 32];
  __shared__ int win[4];
  __shared__ double gw[3][RIBCA_GAUSS_STRIDE];
  __shared__ double s_lohi[2];
  __shared__ int red_cnt[kThreadsR / 32];

  const int j = blockIdx.x;
  if (j >= n_cells) return;
  const int cell = cell_begin + j;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int id = ids[cell];
  if (tid == 0) {
    // utils.py:227-235 with a float half-width: int(max(xm - PW/2, 0)) = max(xm - ceil(PW/2), 0)
    const int4 bb = reinterpret_cast<const int4*>(cbbox)[cell];
    const int half_up = (pw + 1) / 2;
    const int r0 = max(((bb.x + bb.y) >> 1) - half_up, 0), c0 = max(((bb.z + bb.w) >> 1) - half_up, 0);
    win[0] = r0; win[1] = min(r0 + pw, H); win[2] = c0; win[3] = min(c0 + pw, W);
    if (windows) reinterpret_cast<int4*>(windows)[j] = make_int4(win[0], win[1], win[2], win[3]);
  }
  if (tid < 3 * RIBCA_GAUSS_STRIDE) gw[tid / RIBCA_GAUSS_STRIDE][tid % RIBCA_GAUSS_STRIDE] = prm.g[tid / RIBCA_GAUSS_STRIDE][tid % RIBCA_GAUSS_STRIDE];
  __syncthreads();
  const int r0 = win[0], c0 = win[2], wh = win[1] - win[0], ww = win[3] - win[2];

  for (int p = tid; p < n; p += kThreadsR) {
    const int y = p / pw, x = p - y * pw;
    int lab = 0;
    if (y < wh && x < ww) lab = __ldg(mask + (long long)(r0 + y) * W + (c0 + x));
    mk[p] = (unsigned char)((lab == id ? 1 : 0) | (lab > 0 ? 2 : 0));
    s[p] = lab == id ? 1.0f : 0.0f;
  }
  __syncthreads();
  // s = m + sum_j [ d_j + sum_{i < j-1} G_{1+i}(d_j) ]   in the reference's order (utils.py:259-266)
  for (int dj = 1; dj <= 4; ++dj) {
    for (int p = tid; p < n; p += kThreadsR) {
      const int y = p / pw, x = p - y * pw;
      unsigned char hit = 0;
      for (int dy = -dj; dy <= dj && !hit; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= pw) continue;
        const int hwid = disk_halfwidth(dj, dy < 0 ? -dy : dy);
        for (int dx = -hwid; dx <= hwid; ++dx) {
          const int xx = x + dx;
          if (xx >= 0 && xx < pw && (mk[yy * pw + xx] & 1)) { hit = 1; break; }
        }
      }
      dcur[p] = hit;
      s[p] = __fadd_rn(s[p], (float)hit);
    }
    __syncthreads();
    for (int i = 0; i < dj - 1; ++i) add_gaussian_bytes(dcur, pw, gw[i], 4 * (i + 1), tmp, s);
  }
  float mx = 0.0f;
  for (int p = tid; p < n; p += kThreadsR) {
    const float v = __fdiv_rn(s[p], 11.0f);
    s[p] = v;
    mx = fmaxf(mx, __fadd_rn(v, 1e-6f));
  }
  mx = warp_max(mx);
  if (lane == 0) red_f[warp] = mx;
  __syncthreads();
  mx = red_f[0];
#pragma unroll
  for (int k = 1; k < kThreadsR / 32; ++k) mx = fmaxf(mx, red_f[k]);
  for (int p = tid; p < n; p += kThreadsR) s[p] = __fdiv_rn(s[p], mx);
  __syncthreads();

  // ---- pass A over every image channel: range of the float64 patch (for the final clip) and mean intensity ----
  const long long plane = (long long)H * W;
  {
    int cnt = 0;
    for (int p = tid; p < n; p += kThreadsR) cnt += (mk[p] >> 1) & 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) red_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int k = 0; k < kThreadsR / 32; ++k) tot += red_cnt[k];
      red_cnt[0] = tot;
    }
    __syncthreads();
  }
  const int n_lab = red_cnt[0];
  double lo = INFINITY, hi = -INFINITY;
  for (int c = 0; c < C_img; ++c) {
    const float mn = __ldg(min_val + c);
    const float* ip = img + c * plane;
    double acc = 0.0;
    for (int p = tid; p < n; p += kThreadsR) {
      const int y = p / pw, x = p - y * pw;
      float z = 0.0f;
      if (y < wh && x < ww) z = __fsub_rn(__ldg(ip + (long long)(r0 + y) * W + (c0 + x)), mn);
      const double v = __dadd_rn(__dmul_rn((double)z, (double)s[p]), (double)mn);
      lo = fmin(lo, v); hi = fmax(hi, v);
      if (mk[p] & 2) acc += v;
    }
    if (avg_int) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) red_sum[warp] = acc;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < kThreadsR / 32; ++k) t += red_sum[k];
        avg_int[(long long)j * C_img + c] = t / (double)n_lab;
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
  __syncthreads();
  if (tid == 0) {
    double a = red_lo[0], b = red_hi[0];
    for (int k = 1; k < kThreadsR / 32; ++k) { a = fmin(a, red_lo[k]); b = fmax(b, red_hi[k]); }
    s_lohi[0] = a; s_lohi[1] = b;
  }
  __syncthreads();
  lo = s_lohi[0]; hi = s_lohi[1];

  // ---- pass B: resize every requested channel --------------------------------------------------------------
  for (int pnl = 0; pnl < prm.n_panels; ++pnl) {
    const int nch = prm.n_ch[pnl];
    float* dst = prm.out[pnl] + (long long)j * nch * PP;
    for (int k = 0; k < nch; ++k) {
      const int src = prm.src[pnl][k];
      float* o = dst + (long long)k * PP;
      if (src < 0) {
        for (int p = tid; p < PP; p += kThreadsR) o[p] = -1.0f;
        continue;
      }
      const float mn = __ldg(min_val + src);
      const float* ip = img + src * plane;
      for (int p = tid; p < n; p += kThreadsR) {
        const int y = p / pw, x = p - y * pw;
        float z = 0.0f;
        if (y < wh && x < ww) z = __fsub_rn(__ldg(ip + (long long)(r0 + y) * W + (c0 + x)), mn);
        val[p] = __dadd_rn(__dmul_rn((double)z, (double)s[p]), (double)mn);
      }
      __syncthreads();
      if (rz.r_aa > 0) {
        const int r = rz.r_aa;
        for (int p = tid; p < n; p += kThreadsR) {            // axis 0 of the image plane ('mirror')
          const int y = p / pw, x = p - y * pw;
          double acc = __dmul_rn(val[p], rz.w_aa[0]);
          for (int t = r; t >= 1; --t)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(val[mirror_index(y - t, pw) * pw + x], val[mirror_index(y + t, pw) * pw + x]), rz.w_aa[t]));
          tmp[p] = acc;
        }
        __syncthreads();
        for (int p = tid; p < n; p += kThreadsR) {            // axis 1
          const int y = p / pw, x = p - y * pw;
          const double* row = tmp + y * pw;
          double acc = __dmul_rn(row[x], rz.w_aa[0]);
          for (int t = r; t >= 1; --t)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(row[mirror_index(x - t, pw)], row[mirror_index(x + t, pw)]), rz.w_aa[t]));
          val[p] = acc;
        }
        __syncthreads();
      }
      for (int p = tid; p < PP; p += kThreadsR) {
        const int oy = p / P, ox = p - oy * P;
        const double v = val[rz.src[oy] * pw + rz.src[ox]];
        o[p] = (float)fmin(fmax(v, lo), hi);
      }
      __syncthreads();
    }
  }
}

__global__ void channel_min_kernel(const float* __restrict__ img, long long hw, float* min_val) {
  // grid = (blocks_per_channel, C); min over a channel via ordered-int atomicMin on the float bits
  const int c = blockIdx.y;
  const float* p = img + c * hw;
  float m = INFINITY;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += stride) m = fminf(m, __ldg(p + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) {
    // order-preserving map float -> int
    int bits = __float_as_int(m);
    int key = bits >= 0 ? bits : bits ^ 0x7fffffff;
    atomicMin(reinterpret_cast<int*>(min_val) + c, key);
  }
}
__global__ void channel_min_init_kernel(float* min_val, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) reinterpret_cast<int*>(min_val)[c] = INT32_MAX;
}
__global__ void channel_min_finish_kernel(float* min_val, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    int key = reinterpret_cast<int*>(min_val)[c];
    int bits = key >= 0 ? key : key ^ 0x7fffffff;
    min_val[c] = __int_as_float(bits);
  }
}

}  // namespace ribca

using namespace ribca;

extern "C" {

int ribca_channel_min(const float* img, int C, long long hw, float* min_val, ribca_stream_t stream) {
  RIBCA_REQUIRE(img && min_val && C > 0 && hw > 0, "ribca_channel_min: bad arguments");
  cudaStream_t st = as_stream(stream);
  channel_min_init_kernel<<<(C + 63) / 64, 64, 0, st>>>(min_val, C);
  RIBCA_LAUNCH_CHECK("channel_min_init_kernel");
  int bx = (int)std::min<long long>((hw + 1023) / 1024, (long long)num_sms() * 4);
  channel_min_kernel<<<dim3(bx, C), 256, 0, st>>>(img, hw, min_val);
  RIBCA_LAUNCH_CHECK("channel_min_kernel");
  channel_min_finish_kernel<<<(C + 63) / 64, 64, 0, st>>>(min_val, C);
  RIBCA_LAUNCH_CHECK("channel_min_finish_kernel");
  return RIBCA_OK;
}

int ribca_build_patches(const float* img, const int32_t* mask, int C_img, int H, int W,
                        const float* min_val, const int32_t* ids, const int32_t* cbbox,
                        int cell_begin, int n_cells, int n_panels, const int* h_n_ch,
                        const int* h_chan_index, float* const* h_out, const double* h_gauss,
                        double* avg_int, int32_t* windows, ribca_stream_t stream) {
  RIBCA_REQUIRE(img && mask && min_val && ids && cbbox && h_gauss, "ribca_build_patches: null pointer");
  RIBCA_REQUIRE(C_img > 0 && H > 0 && W > 0 && cell_begin >= 0 && n_cells >= 0, "ribca_build_patches: bad shape");
  RIBCA_REQUIRE(n_panels >= 0 && n_panels <= RIBCA_MAX_PANELS, "ribca_build_patches: n_panels=%d out of range", n_panels);
  RIBCA_REQUIRE(n_panels == 0 || (h_n_ch && h_chan_index && h_out), "ribca_build_patches: null panel arrays");
  if (n_cells == 0) return RIBCA_OK;
  PatchParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.n_panels = n_panels;
  for (int p = 0; p < n_panels; ++p) {
    const int nch = h_n_ch[p];
    RIBCA_REQUIRE(nch > 0 && nch <= RIBCA_MAX_PANEL_CH, "ribca_build_patches: panel %d has %d channels", p, nch);
    RIBCA_REQUIRE(h_out[p] != nullptr, "ribca_build_patches: panel %d output is null", p);
    prm.out[p] = h_out[p];
    prm.n_ch[p] = nch;
    bool blank_used = false;     // only the first -1 is a blank plane (quirk Q3)
    for (int k = 0; k < nch; ++k) {
      int idx = h_chan_index[p * RIBCA_MAX_PANEL_CH + k];
      if (idx == -1) {
        if (!blank_used) { prm.src[p][k] = -1; blank_used = true; }
        else prm.src[p][k] = C_img - 1;
      } else {
        RIBCA_REQUIRE(idx >= 0 && idx < C_img, "ribca_build_patches: channel index %d outside [0,%d)", idx, C_img);
        prm.src[p][k] = idx;
      }
    }
  }
  for (int sgm = 0; sgm < 3; ++sgm)
    for (int k = 0; k < RIBCA_GAUSS_STRIDE; ++k) prm.g[sgm][k] = h_gauss[sgm * RIBCA_GAUSS_STRIDE + k];
  const bool prof = profiling();
  if (prof) {   // algorithmic bytes: window reads of the used channels + mask, patch writes
    double out_ch = 0;
    for (int p = 0; p < n_panels; ++p) out_ch += h_n_ch[p];
    prof_begin_span(RIBCA_PROF_PATCHES, (double)n_cells * 1600.0 * (4.0 * out_ch + 4.0 + 4.0 * out_ch), as_stream(stream));
  }
  build_patches_kernel<<<n_cells, kThreads, 0, as_stream(stream)>>>(img, mask, C_img, H, W, min_val, ids, cbbox,
                                                                    cell_begin, n_cells, prm, avg_int, windows);
  if (prof) prof_end_span(as_stream(stream));
  RIBCA_LAUNCH_CHECK("build_patches_kernel");
  return RIBCA_OK;
}


int ribca_build_patches_resized(const float* img, const int32_t* mask, int C_img, int H, int W, const float* min_val,
                                const int32_t* ids, const int32_t* cbbox, int cell_begin, int n_cells, int n_panels,
                                const int* h_n_ch, const int* h_chan_index, float* const* h_out, const double* h_gauss,
                                int patch_edge, const int* h_src_index, const double* h_w_aa, int r_aa, double* avg_int,
                                int32_t* windows, ribca_stream_t stream) {
  RIBCA_REQUIRE(img && mask && min_val && ids && cbbox && h_gauss && h_src_index, "ribca_build_patches_resized: null pointer");
  RIBCA_REQUIRE(patch_edge >= 8 && patch_edge <= kMaxPW, "ribca_build_patches_resized: patch edge %d outside [8, %d]", patch_edge, kMaxPW);
  RIBCA_REQUIRE(r_aa < 8 && (r_aa <= 0 || h_w_aa), "ribca_build_patches_resized: bad anti-alias kernel");
  RIBCA_REQUIRE(n_panels >= 0 && n_panels <= RIBCA_MAX_PANELS && C_img > 0, "ribca_build_patches_resized: bad panel count");
  if (n_cells <= 0) return RIBCA_OK;
  PatchParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.n_panels = n_panels;
  for (int p = 0; p < n_panels; ++p) {
    const int nch = h_n_ch[p];
    RIBCA_REQUIRE(nch > 0 && nch <= RIBCA_MAX_PANEL_CH && h_out[p], "ribca_build_patches_resized: bad panel %d", p);
    prm.out[p] = h_out[p];
    prm.n_ch[p] = nch;
    bool blank_used = false;
    for (int k = 0; k < nch; ++k) {
      int idx = h_chan_index[p * RIBCA_MAX_PANEL_CH + k];
      if (idx == -1) {
        if (!blank_used) { prm.src[p][k] = -1; blank_used = true; }
        else prm.src[p][k] = C_img - 1;
      } else {
        RIBCA_REQUIRE(idx >= 0 && idx < C_img, "ribca_build_patches_resized: channel index %d outside [0,%d)", idx, C_img);
        prm.src[p][k] = idx;
      }
    }
  }
  for (int sgm = 0; sgm < 3; ++sgm)
    for (int k = 0; k < RIBCA_GAUSS_STRIDE; ++k) prm.g[sgm][k] = h_gauss[sgm * RIBCA_GAUSS_STRIDE + k];
  ResizeParams rz;
  memset(&rz, 0, sizeof(rz));
  rz.pw = patch_edge;
  rz.r_aa = r_aa;
  for (int i = 0; i < P; ++i) {
    RIBCA_REQUIRE(h_src_index[i] >= 0 && h_src_index[i] < patch_edge, "ribca_build_patches_resized: source index out of range");
    rz.src[i] = h_src_index[i];
  }
  for (int k = 0; k <= r_aa && k < 8; ++k) rz.w_aa[k] = h_w_aa[k];
  const size_t n = (size_t)patch_edge * patch_edge;
  const size_t smem = n * (8 + 8 + 4 + 1 + 1) + 16;
  RIBCA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(build_patches_resized_kernel), (int)((int)((size_t)kMaxPW * kMaxPW * 22 + 16)), "cudaFuncSetAttribute(build_patches_resized_kernel)"));
  build_patches_resized_kernel<<<n_cells, kThreadsR, smem, as_stream(stream)>>>(img, mask, C_img, H, W, min_val, ids, cbbox, cell_begin,
                                                                                n_cells, prm, rz, avg_int, windows);
  RIBCA_LAUNCH_CHECK("build_patches_resized_kernel");
  return RIBCA_OK;
}

}  // extern "C"
