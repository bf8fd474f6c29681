// Stage 5: vote merge, "Others" threshold and per-type counts in one pass over the softmax tables.
// Replaces Annotator.merge_by_voting + get_void_vote (reference cta/model.py:481-636,
// cta/utils.py:143-146).  All comparisons are float32 (numpy >= 2 rounds the Python-float
// thresholds to float32 before comparing with the np.float32 votes).
//
//   one model : best = first argmax in class order (incl. "Others"); thr = ctc[best] > 0 ? ctc[best]
//               : confidence; best != Others and p < thr -> ("Others", -1) else (best, p)
//   two models: vote[type] = the probability of the (single) model that predicts that type, 0 for
//               types no model predicts; best = first maximum in get_void_vote() key order;
//               thr = ctc[best] < 0 ? min(others_0, others_1, confidence) : ctc[best];
//               vote < thr -> ("Others", -1) else (best, vote)
// HBM-bound: 4 * (classes0 + classes1) bytes read and 5 bytes written per cell.
#include "common.cuh"

namespace ribca {

constexpr int kTypes = 18;
constexpr int kOthers = 17;
constexpr int kMaxClasses = 16;

struct MergeParams {
  int classes[2];
  int type_of_class[2][kMaxClasses];
  int order[kTypes];          // order[i] = global type at position i of the vote key order (17 entries)
  float type_thresh[kTypes];
  float confidence;
};

// Decision margin (the guard of the exact-label re-evaluation, pipeline.refine_labels): the smallest perturbation of
// the probabilities that could change the OUTCOME of the cell = (label, "was re-labelled Others").  With the winner k
// (value v_k, threshold thr_k) and every other candidate j that could win instead:
//   m_thr  = |v_k - thr_k|                                   (the winner crossing its threshold; not for the class "Others")
//   m_j    = v_k - v_j                    if candidate j would give another outcome at its current value,
//            max(v_k - v_j, |v_j - thr_j|) otherwise           (j must both overtake k and cross its own threshold)
//   margin = min(m_thr, min_j m_j)
// A cell whose margin exceeds twice the error bound of the probabilities has the same outcome in exact arithmetic.
__device__ __forceinline__ int outcome_of(int type, float v, float thr, bool single) {
  // single model: "Others" is a class of its own and is never re-labelled; two models: every winner is thresholded
  if (single) return (type != kOthers && v < thr) ? (kOthers | 32) : type;
  return v < thr ? (kOthers | 32) : type;
}

__global__ void __launch_bounds__(256)
merge_votes_kernel(const float* __restrict__ p0, const float* __restrict__ p1, int n,
                   const __grid_constant__ MergeParams prm, uint8_t* __restrict__ label,
                   float* __restrict__ conf, long long* __restrict__ counts, float* __restrict__ margin) {
  __shared__ int hist[kTypes];
  if (threadIdx.x < kTypes) hist[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int best;
    float value;
    float mg = INFINITY;
    if (p1 == nullptr) {
      const float* row = p0 + (long long)i * prm.classes[0];
      int k = 0;
      float pk = row[0];
      for (int c = 1; c < prm.classes[0]; ++c) {
        const float v = row[c];
        if (v > pk) { pk = v; k = c; }
      }
      best = prm.type_of_class[0][k];
      const float t = prm.type_thresh[best];
      const float thr = t > 0.0f ? t : prm.confidence;
      value = pk;
      if (margin) {
        const int out_k = outcome_of(best, pk, thr, true);
        if (best != kOthers) mg = fabsf(pk - thr);
        for (int c = 0; c < prm.classes[0]; ++c) {
          if (c == k) continue;
          const int tj = prm.type_of_class[0][c];
          const float tt = prm.type_thresh[tj];
          const float thr_j = tt > 0.0f ? tt : prm.confidence;
          const float vj = row[c];
          float mj = pk - vj;
          if (outcome_of(tj, vj, thr_j, true) == out_k) mj = tj == kOthers ? INFINITY : fmaxf(mj, fabsf(vj - thr_j));
          mg = fminf(mg, mj);
        }
      }
      if (best != kOthers && pk < thr) { best = kOthers; value = -1.0f; }
    } else {
      float vote[kTypes];
#pragma unroll
      for (int t = 0; t < kTypes; ++t) vote[t] = 0.0f;
      float others[2] = {0.0f, 0.0f};
      const float* rows[2] = {p0 + (long long)i * prm.classes[0], p1 + (long long)i * prm.classes[1]};
      for (int m = 0; m < 2; ++m)
        for (int c = 0; c < prm.classes[m]; ++c) {
          const int t = prm.type_of_class[m][c];
          const float v = rows[m][c];
          if (t == kOthers) others[m] = v;
          else vote[t] = __fadd_rn(vote[t], v);
        }
      best = prm.order[0];
      float bv = vote[best];
      for (int q = 1; q < kTypes - 1; ++q) {
        const int t = prm.order[q];
        if (vote[t] > bv) { bv = vote[t]; best = t; }
      }
      const float t = prm.type_thresh[best];
      // Python min(o1, o2, confidence): first minimal element, float32 comparisons
      float thr_min = others[0];
      if (others[1] < thr_min) thr_min = others[1];
      if (prm.confidence < thr_min) thr_min = prm.confidence;
      const float thr = !(t < 0.0f) ? t : thr_min;
      value = bv;
      if (margin) {
        const int out_k = outcome_of(best, bv, thr, false);
        mg = fabsf(bv - thr);
        for (int q = 0; q < kTypes - 1; ++q) {
          const int tj = prm.order[q];
          if (tj == best) continue;
          const float tt = prm.type_thresh[tj];
          const float thr_j = !(tt < 0.0f) ? tt : thr_min;
          const float vj = vote[tj];
          float mj = bv - vj;
          if (outcome_of(tj, vj, thr_j, false) == out_k) mj = fmaxf(mj, fabsf(vj - thr_j));
          mg = fminf(mg, mj);
        }
      }
      if (bv < thr) { best = kOthers; value = -1.0f; }
    }
    label[i] = (uint8_t)best;
    conf[i] = value;
    if (margin) margin[i] = mg;
    atomicAdd(&hist[best], 1);
  }
  __syncthreads();
  if (counts && threadIdx.x < kTypes && hist[threadIdx.x])
    atomicAdd(reinterpret_cast<unsigned long long*>(counts) + threadIdx.x, (unsigned long long)hist[threadIdx.x]);
}

// per-pixel maps from per-cell values: out[p] = mask[p] > 0 ? value[id_to_index[mask[p]]] : 0
// (the gather behind Annotator.colorize, reference cta/model.py:806-858, without per-cell pixel lists)
__global__ void __launch_bounds__(256)
paint_cells_kernel(const int32_t* __restrict__ mask, long long n, const int32_t* __restrict__ id_to_index, int max_id,
                   const uint8_t* __restrict__ value, int channels, uint8_t* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int id = __ldg(mask + i);
    int k = -1;
    if (id > 0 && id <= max_id) k = __ldg(id_to_index + id);
    for (int c = 0; c < channels; ++c) out[i * channels + c] = k >= 0 ? __ldg(value + (long long)k * channels + c) : (uint8_t)0;
  }
}

}  // namespace ribca

using namespace ribca;

extern "C" int ribca_paint_cells(const int32_t* mask, long long n_pixels, const int32_t* id_to_index, int max_id,
                                 const uint8_t* cell_value, int channels, uint8_t* out, ribca_stream_t stream) {
  RIBCA_REQUIRE(mask && id_to_index && cell_value && out && n_pixels > 0 && channels > 0 && channels <= 4,
                "ribca_paint_cells: bad arguments");
  const int blocks = (int)std::min<long long>((n_pixels + 255) / 256, (long long)num_sms() * 16);
  paint_cells_kernel<<<blocks, 256, 0, as_stream(stream)>>>(mask, n_pixels, id_to_index, max_id, cell_value, channels, out);
  RIBCA_LAUNCH_CHECK("paint_cells_kernel");
  return RIBCA_OK;
}

extern "C" int ribca_merge_votes(const float* probs0, int classes0, const int* h_type_of_class0,
                                 const float* probs1, int classes1, const int* h_type_of_class1, int n_cells,
                                 const int* h_vote_rank, const float* h_type_thresh, float confidence,
                                 uint8_t* label, float* conf, long long* counts, float* margin, ribca_stream_t stream) {
  RIBCA_REQUIRE(probs0 && h_type_of_class0 && h_vote_rank && h_type_thresh && label && conf,
                "ribca_merge_votes: null pointer");
  RIBCA_REQUIRE(classes0 > 0 && classes0 <= kMaxClasses, "ribca_merge_votes: classes0=%d", classes0);
  RIBCA_REQUIRE(!probs1 || (classes1 > 0 && classes1 <= kMaxClasses && h_type_of_class1), "ribca_merge_votes: bad second model");
  if (n_cells <= 0) return RIBCA_OK;
  MergeParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.classes[0] = classes0;
  prm.classes[1] = probs1 ? classes1 : 0;
  for (int c = 0; c < classes0; ++c) {
    RIBCA_REQUIRE(h_type_of_class0[c] >= 0 && h_type_of_class0[c] < kTypes, "ribca_merge_votes: bad type index");
    prm.type_of_class[0][c] = h_type_of_class0[c];
  }
  for (int c = 0; c < prm.classes[1]; ++c) {
    RIBCA_REQUIRE(h_type_of_class1[c] >= 0 && h_type_of_class1[c] < kTypes, "ribca_merge_votes: bad type index");
    prm.type_of_class[1][c] = h_type_of_class1[c];
  }
  for (int t = 0; t < kTypes; ++t) prm.order[t] = kOthers;
  for (int t = 0; t < kTypes - 1; ++t) {
    const int rank = h_vote_rank[t];
    RIBCA_REQUIRE(rank >= 0 && rank < kTypes - 1, "ribca_merge_votes: vote rank %d of type %d out of range", rank, t);
    prm.order[rank] = t;
  }
  for (int t = 0; t < kTypes; ++t) prm.type_thresh[t] = h_type_thresh[t];
  prm.confidence = confidence;
  const bool prof = profiling();
  if (prof) prof_begin_span(RIBCA_PROF_MERGE, (double)n_cells * (4.0 * (prm.classes[0] + prm.classes[1]) + 5.0), as_stream(stream));
  merge_votes_kernel<<<(n_cells + 255) / 256, 256, 0, as_stream(stream)>>>(probs0, probs1, n_cells, prm, label, conf, counts, margin);
  if (prof) prof_end_span(as_stream(stream));
  RIBCA_LAUNCH_CHECK("merge_votes_kernel");
  return RIBCA_OK;
}
