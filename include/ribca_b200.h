/*
 * ribca_b200.h - C ABI of libribca_b200.so: RIBCA's per-cell annotation hot path on B200 (sm_100a).
 *
 * The reference (sun-huangqingbo/multiplexed-image-annotator) is pure Python and has no FFI of its
 * own; its drop-in boundary is the Python API.  This library sits UNDER the
 * Python classes that mirror that API and replaces, one entry point per stage, the reference
 * functions cited beside each declaration (paths relative to the reference root,
 * cta/ = src/multiplexed_image_annotator/cell_type_annotation/).
 *
 * Conventions
 *   - every function returns 0 on success or a negative RIBCA_E* code; ribca_last_error() returns
 *     a thread-local message for the last failure.  No C++ exception crosses the boundary.
 *   - all buffers are caller-allocated; pointers are DEVICE pointers unless the parameter name
 *     starts with h_ (small host arrays that are copied into kernel parameters).
 *   - no hidden allocation: scratch is an explicit `workspace` sized by the matching
 *     *_workspace_bytes() query.  Calls are asynchronous on `stream` (a cudaStream_t) and
 *     re-entrant per stream.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     RIBCA_ECUDA.
 */
#ifndef RIBCA_B200_H
#define RIBCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIBCA_OK 0
#define RIBCA_EINVAL (-1)   /* bad argument                           */
#define RIBCA_ECUDA (-2)    /* CUDA runtime / driver error            */
#define RIBCA_EWORKSPACE (-3) /* workspace too small                  */
#define RIBCA_EUNSUPPORTED (-4)

#define RIBCA_PATCH 40         /* model input patch edge (cta/preprocess.py:76,99)      */
#define RIBCA_MAX_TAPS 128     /* largest Gaussian radius accepted by the FIR kernels   */
#define RIBCA_MAX_PANEL_CH 16  /* largest panel (immune_full has 15 markers)            */
#define RIBCA_MAX_PANELS 3     /* one immune panel + structure + nerve per predict()    */
#define RIBCA_MAX_TYPES 18     /* cell types of cta/model.py:97-99                       */

typedef void* ribca_stream_t;  /* cudaStream_t */

enum ribca_dtype { RIBCA_U8 = 0, RIBCA_U16 = 1, RIBCA_F32 = 2, RIBCA_I32 = 3 };

/* operand precision of the tensor-core contractions (stage 4) */
enum ribca_precision {
  RIBCA_BF16X3 = 0, /* split-bf16, 3 tcgen05 passes (hi*hi + lo*hi + hi*lo), fp32 accumulate: parity mode */
  RIBCA_BF16X1 = 1, /* single bf16 pass: throughput mode (outside the 1e-3 probability tolerance)          */
  RIBCA_SIMT_FP32 = 2, /* same contraction on the FP32 pipe (debug cross-check, no tensor cores)            */
  RIBCA_F16F8 = 3   /* fp16 main pass + one e4m3 pass over a doubled K axis that carries both first-order
                       correction terms (a_lo*w_hi + a_hi*w_lo): two passes' worth of tensor time, relative
                       error ~2^-15 per product (measured max |dprob| 1.3e-4): default                    */
  , RIBCA_FP32 = 4  /* plain fp32 operands and fp32 FMA accumulation on the FP32 pipe (the reference's own arithmetic,
                       cta/model.py:397-406): last level of the exact-label re-evaluation; weights = RIBCA_PLANES_F32 */
};
/* operand formats of the packed GEMM weights: two 16-bit planes (see csrc/common.cuh) or plain fp32 matrices */
enum ribca_plane_format { RIBCA_PLANES_BF16 = 0, RIBCA_PLANES_F16F8 = 1, RIBCA_PLANES_F32 = 2 };

const char* ribca_last_error(void);
int ribca_version(void);
/* number of kernels this library has launched in the calling process (bench.py `gpu_launches`) */
long long ribca_launch_count(void);
/* Two-way interleave of ribca_vit_forward (opt-in: RIBCA_INTERLEAVE=1 in the environment or on = 1 here; measured
 * neutral on the power-capped B200s of this pool, profiles/r02_interleave.md): the cells of one call are split in two halves that run the same kernel sequence on the caller's stream and on a
 * library-owned side stream (fork / join by events, so the caller's stream ordering is unchanged).  The HBM-bound
 * kernels of one half (LayerNorm, im2col, head) then share the SMs with the tensor-core GEMM of the other half, which
 * leaves the registers and all but 224 KB of shared memory to them; every row is computed by the same kernels in the same
 * order, so the probabilities are bit-identical to the serial schedule.  No reference counterpart (the reference runs
 * one torch op at a time, cta/model.py:397-406). */
int ribca_set_interleave(int on);

/* Optional per-kernel-class device timing for roofline reports: between ribca_profile_begin() and
 * ribca_profile_end() every launch of the classes below is bracketed by CUDA events on its stream.
 * ribca_profile_end fills, per class, the summed kernel time (ms), the launch count and the
 * algorithmic work (FLOP for GEMM / attention, bytes for the patch builder). */
#define RIBCA_PROF_GEMM 0
#define RIBCA_PROF_ATTENTION 1
#define RIBCA_PROF_PATCHES 2
#define RIBCA_PROF_NORMALIZE 3   /* whole ribca_normalize call; work = C*H*W*(2*sizeof(in)+4) bytes */
#define RIBCA_PROF_CELLSTATS 4   /* ribca_cell_stats;           work = H*W*4 bytes                    */
#define RIBCA_PROF_LAYERNORM 5   /* work = M*D*8 bytes                                                */
#define RIBCA_PROF_MERGE 6       /* work = n*(4*classes+5) bytes                                      */
#define RIBCA_PROF_GEMM_F32 7    /* sgemm_nt_kernel of the RIBCA_FP32 re-evaluation; work = 2*M*N*K FLOP */
#define RIBCA_PROF_CLASSES 8
int ribca_profile_begin(void);
int ribca_profile_end(double* ms, long long* launches, double* work, int n_classes);

/* ---------------------------------------------------------------------------------------------
 * Stage 1 - image normalisation.      replaces ImageProcessor._normalize  cta/preprocess.py:214-239
 *
 * Per channel: x = f32(img); bg = G_sigma20(x) (scipy gaussian_filter: axis 0 then axis 1, mode
 * reflect, float64 accumulate in scipy's tap order, float32 store) ; bg = min(bg, 125);
 * x = max(x - bg, 0); optional blur G_blur(x); all-non-positive channel -> -1; t = percentile
 * (numpy 'linear' between the order statistics k_lo, k_hi with weight gamma, all float32);
 * clip to t when t > 20; out = 2 * x / max(25, max x) - 1.
 * The tap weights and the order-statistic plan are computed by the host exactly as scipy / numpy
 * do (so results are bit-identical to the reference) and passed in:
 *   h_w_bg[0..r_bg]     half kernel, h_w_bg[j] = weight at distance j from the centre
 *   h_w_blur[0..r_blur] same for the blur; r_blur < 0 disables the blur
 *   k_lo, k_hi, gamma   np.percentile plan for n = H*W (see host `percentile_plan`)
 * chan_stats (optional, device, C x 4 floats): [percentile, max after clip, all_non_positive, min(out)].
 */
size_t ribca_normalize_workspace_bytes(int C, int H, int W);
int ribca_normalize(const void* img, int dtype, int C, int H, int W,
                    const double* h_w_bg, int r_bg, const double* h_w_blur, int r_blur,
                    long long k_lo, long long k_hi, float gamma,
                    float* out, float* chan_stats, void* workspace, size_t workspace_bytes,
                    ribca_stream_t stream);

/* per-channel minimum of a float32 stack (ImageProcessor._move_image_range, cta/preprocess.py:153-157) */
int ribca_channel_min(const float* img, int C, long long hw, float* min_val, ribca_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 - label mask -> per-cell statistics.
 *                 replaces ImageProcessor._cell_pos_dict + process_chunk
 *                          cta/preprocess.py:159-211, cta/utils.py:272-290
 * and the reductions every consumer applies to the pixel lists (bbox: cta/utils.py:227,232;
 * centroid: cta/model.py:785-786; area).  Integer, bit-exact.
 *
 * ribca_mask_minmax: out2[0] = min label, out2[1] = max label (device int32[2]).
 * ribca_cell_stats:  dense tables indexed by label 0..max_id (label 0 = background is skipped):
 *     bbox[id] = {rmin, rmax, cmin, cmax}, sums[id] = {sum of rows, sum of cols}, count[id].
 *     The call initialises the tables itself.
 * ribca_compact_cells: ascending list of the labels with count > 0 and their rows of the tables
 *     -> ids[n], cbbox[n][4], csums[n][2], ccount[n]; *n_cells (device int32).
 */
int ribca_mask_minmax(const int32_t* mask, long long n, int32_t* out2, ribca_stream_t stream);
int ribca_cell_stats(const int32_t* mask, int H, int W, int max_id, int32_t* bbox,
                     unsigned long long* sums, int32_t* count, ribca_stream_t stream);
size_t ribca_compact_workspace_bytes(int max_id);
int ribca_compact_cells(const int32_t* bbox, const unsigned long long* sums, const int32_t* count,
                        int max_id, int32_t* ids, int32_t* cbbox, unsigned long long* csums,
                        int32_t* ccount, int32_t* id_to_index, int32_t* n_cells, void* workspace,
                        size_t workspace_bytes, ribca_stream_t stream);

/* CSR pixel lists = the reference's cell_pos_dict itself (cta/preprocess.py:159-181): for compact cell j
 * (label ids[j], bbox cbbox[j]) the rows / cols of its pixels in raster order are written to
 * rows[offsets[j] .. offsets[j+1]) and cols[...]; offsets = exclusive prefix sum of ccount (n_cells + 1
 * int64, computed by the caller).  Integer, bit-exact. */
int ribca_cell_pixels(const int32_t* mask, int H, int W, const int32_t* ids, const int32_t* cbbox,
                      const long long* offsets, int n_cells, int32_t* rows, int32_t* cols,
                      ribca_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 - per-cell patch gather with the soft cell mask.
 *      replaces crop_cell + smooth (cta/utils.py:226-270) and ImageProcessor._img2patches
 *      (cta/preprocess.py:76-151) for cell_size = 30 (40 x 40 patch, resize = identity).
 *
 * For cell j (label ids[j], bbox cbbox[j]): window (x0,x1,y0,y1) of cta/utils.py:227-235;
 * s = smooth(mask window == id) in the reference's float32/float64 op order (bit-exact);
 * value[c] = f32( f64(img[c] - min_val[c]) * f64(s) + f64(min_val[c]) ), zero-padded window;
 * each requested panel p gets out[p][j][k] = value[h_chan_index[p][k]] or -1 for the first -1
 * entry (a later -1 selects the last image channel: numpy negative index, quirk Q3).
 * avg_int[j][c] (optional, float64) = mean of the float64 value over window pixels with label > 0.
 * windows[j] (optional) = {x0, x1, y0, y1}.
 *   h_gauss[3][RIBCA_GAUSS_STRIDE]: half kernels of sigma 1, 2, 3 (radius 4, 8, 12) as scipy
 *   computes them (host numpy), h_gauss[s][j] = weight at distance j.
 */
#define RIBCA_GAUSS_STRIDE 16
int ribca_build_patches(const float* img, const int32_t* mask, int C_img, int H, int W,
                        const float* min_val, const int32_t* ids, const int32_t* cbbox,
                        int cell_begin, int n_cells, int n_panels, const int* h_n_ch,
                        const int* h_chan_index /* [n_panels][RIBCA_MAX_PANEL_CH] */,
                        float* const* h_out /* n_panels device pointers, each [n_cells][n_ch][40][40] */,
                        const double* h_gauss, double* avg_int, int32_t* windows,
                        ribca_stream_t stream);

/* The same for cell_size != 30 (cta/preprocess.py:67,78,106): the window edge is patch_edge =
 * int(40 * cell_size / 30) (<= 80), and the float64 patch is resampled to 40 x 40 like
 * skimage.transform.resize(order=0, anti_aliasing=True, preserve_range=True): Gaussian of sigma
 * (patch_edge/40 - 1)/2 ('mirror', half kernel h_w_aa[0..r_aa], r_aa <= 0 = none), nearest source indices
 * h_src_index[40] (scipy zoom grid_mode arithmetic, computed by the host), clip to the patch range. */
int ribca_build_patches_resized(const float* img, const int32_t* mask, int C_img, int H, int W,
                                const float* min_val, const int32_t* ids, const int32_t* cbbox,
                                int cell_begin, int n_cells, int n_panels, const int* h_n_ch,
                                const int* h_chan_index, float* const* h_out, const double* h_gauss,
                                int patch_edge, const int* h_src_index, const double* h_w_aa, int r_aa,
                                double* avg_int, int32_t* windows, ribca_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 4 - networks.  replaces VisionTransformer / vit_* and _predict_cell_types' forward +
 * softmax (cta/model.py:31-88, 397-406) and MaskedAutoencoderViT / MarkerImputer.impute
 * (cta/markerImputer.py:155-232, 294-329).
 *
 * Primitive: C[M,N] (+epilogue) = A[M,K] . W[N,K]^T with split-bf16 operands.
 *   A, W are stored as two bf16 planes {hi, lo} (x = hi + lo), plane stride a_plane / w_plane
 *   elements, both K-major (row-major [rows][K]); fp32 accumulate in TMEM.
 *   epilogue: v = acc + (bias ? bias[col] : 0) + (row_table ? row_table[(row % table_period)*N + col] : 0)
 *     RIBCA_EPI_STORE    out_f32[row*N+col]  = v
 *     RIBCA_EPI_RESIDUAL out_f32[row*N+col] += v
 *     RIBCA_EPI_GELU        out split-bf16 {hi, lo}[row*N+col] = gelu_erf(v), plane stride out_plane
 *     RIBCA_EPI_STORE_SPLIT out split-bf16 {hi, lo}[row*N+col] = v
 */
enum ribca_epilogue { RIBCA_EPI_STORE = 0, RIBCA_EPI_RESIDUAL = 1, RIBCA_EPI_GELU = 2, RIBCA_EPI_STORE_SPLIT = 3,
                      RIBCA_EPI_STORE_LN = 4, RIBCA_EPI_RESIDUAL_LN = 5 /* ribca_gemm_ln only */ };
int ribca_gemm_splitbf16(const void* A, long long a_plane, const void* W, long long w_plane,
                         int M, int N, int K, const float* bias, const float* row_table,
                         int table_period, int epilogue, float* out_f32, void* out_split,
                         long long out_plane, int precision, int w_log2_scale, ribca_stream_t stream);
/* LayerNorm folded into the GEMMs around it (timm Block: x + attn(norm1(x)), x + mlp(norm2(x)), cta/model.py:54-55), so that
 * no LayerNorm kernel reads the residual stream:
 *   producer (RIBCA_EPI_STORE_LN / RIBCA_EPI_RESIDUAL_LN: patch embedding, proj, fc2): stores the fp32 rows (out_f32 = v or
 *     out_f32 += v) and, from the same values, their operand planes (out_split, plane stride out_plane, in the format of
 *     `precision`) and per-row partial sums ln->stats_out[row][slot] = (sum, sum of squares) over the columns of that slot;
 *     ribca_gemm_ln_slots(N, precision) slots are filled, the partition of columns over slots is fixed (bit-reproducible);
 *   consumer (ln->stats_in != NULL; qkv, fc1): A = the planes of the RAW rows, W' = W * diag(gamma) packed as usual,
 *     ln->c1[n] = sum_k W'[n][k], bias[n] = c2[n] = b[n] + sum_k beta[k] W[n][k];  with mean / rstd of the row from the
 *     statistics over K elements:  v = rstd * (acc - mean * c1[col]) + c2[col]  ( = LayerNorm(x) . W^T + b ), then the
 *     epilogue proper (RIBCA_EPI_STORE / GELU / STORE_SPLIT).
 * ln == NULL: exactly ribca_gemm_splitbf16. */
#define RIBCA_LN_SLOTS 8
typedef struct ribca_ln_fold {
  const float* stats_in;   /* [M][RIBCA_LN_SLOTS][2] or NULL */
  const float* c1;         /* [N] (stats_in != NULL) */
  int slots_in;            /* filled slots of stats_in */
  float eps;               /* LayerNorm epsilon */
  float* stats_out;        /* [M][RIBCA_LN_SLOTS][2] (the *_LN epilogues) */
} ribca_ln_fold;
int ribca_gemm_ln(const void* A, long long a_plane, const void* W, long long w_plane,
                  int M, int N, int K, const float* bias, const float* row_table,
                  int table_period, int epilogue, float* out_f32, void* out_split,
                  long long out_plane, int precision, int w_log2_scale, const ribca_ln_fold* ln,
                  ribca_stream_t stream);
int ribca_gemm_ln_slots(int N, int precision);
/* With precision RIBCA_F16F8 the operands are RIBCA_PLANES_F16F8 planes: plane 0 = fp16, plane 1 = two
 * e4m3 per element; W is packed in the W role with scale 2^t (ribca_split_planes) and w_log2_scale = t + 8
 * (the accumulator is multiplied by 2^-w_log2_scale).  A RIBCA_EPI_GELU output is written in the same
 * format (it feeds the next GEMM); a RIBCA_EPI_STORE_SPLIT output stays bf16 {hi, lo} (it feeds attention).
 * w_log2_scale is ignored by the other precisions. */
/* x (fp32, n) -> bf16 planes hi / lo */
int ribca_split_bf16(const float* x, long long n, void* hi, void* lo, ribca_stream_t stream);
/* x (fp32, n even) -> two 16-bit planes in `format` (ribca_plane_format); w_role != 0 packs a weight matrix
 * scaled by 2^log2_scale (RIBCA_PLANES_F16F8 only; see csrc/common.cuh for the exact encoding) */
int ribca_split_planes(const float* x, long long n, int format, int w_role, int log2_scale,
                       void* plane0, void* plane1, ribca_stream_t stream);

/* row LayerNorm(eps) of x[M][D] -> split-bf16 planes (plane stride out_plane elements) */
int ribca_layernorm_split(const float* x, int M, int D, const float* gamma, const float* beta,
                          float eps, void* out_split, long long out_plane, int format,
                          ribca_stream_t stream);
/* multi-head self-attention over qkv[cells][tokens][3][heads][hd] (fp32) -> split-bf16 [M][D];
 * FP32-pipe kernel, used for the short sequences of the imputer (tokens <= 32) */
int ribca_attention(const float* qkv, int cells, int tokens, int heads, int head_dim,
                    void* out_split, long long out_plane, int format, ribca_stream_t stream);
/* the same on the tensor cores (tcgen05, split-bf16 passes) for tokens <= 112: qkv_split is
 * [2][M][3][heads][hdp] bf16 with hdp = head_dim rounded up to 16 and exact zeros in the padding */
int ribca_attention_tc(const void* qkv_split, long long qkv_plane, int cells, int tokens, int heads,
                       int head_dim, void* out_split, long long out_plane, int format,
                       ribca_stream_t stream);

/* Classifier: device-resident weights in the layout produced by the host packer
 * (multiplexed_image_annotator_b200/engine.py: pack_vit); all offsets in `desc` are element
 * offsets into `wf32` (fp32 params) or `wsplit` (bf16 planes, plane stride desc->split_plane). */
typedef struct ribca_block_desc {
  long long ln1_g, ln1_b, ln2_g, ln2_b;           /* wf32 */
  long long qkv_b, proj_b, fc1_b, fc2_b;           /* wf32 */
  long long qkv_w, proj_w, fc1_w, fc2_w;           /* wsplit */
  /* LayerNorm-folded packing (ribca_vit_desc.ln_folded): qkv_w / fc1_w hold W * diag(gamma); wf32 vectors
   * c1[n] = sum_k (W gamma)[n][k] and c2[n] = b[n] + sum_k beta[k] W[n][k] (qkv ones head-padded like qkv_b) */
  long long qkv_c1, qkv_c2, fc1_c1, fc1_c2;
} ribca_block_desc;
/* qkv_w / qkv_b are stored head-padded: [3][heads][hdp][D] and [3][heads][hdp] with hdp = head_dim
 * rounded up to a multiple of 16 and zero rows in the padding (hdp == head_dim for 32 / 48 / 64). */

typedef struct ribca_vit_desc {
  int dim, heads, depth, in_chans, classes, tokens;  /* tokens = 101 */
  int plane_format;                                  /* ribca_plane_format of wsplit */
  int w_log2_scale;                                  /* RIBCA_PLANES_F16F8: t + 8 (weights packed with scale 2^t) */
  int ln_folded;                                     /* bit 0: norm1 is folded into qkv (qkv_w packed as W diag(gamma), qkv_c1 / qkv_c2 valid), bit 1:
                                                        norm2 into fc1 (ribca_ln_fold); tensor-core formats only */
  int reserved;
  long long split_plane;                             /* elements between the hi and lo plane */
  long long embed_w;                                 /* wsplit [dim][16*in_chans] */
  long long embed_table;                             /* wf32 [tokens][dim]: row0 = cls+pos0, row t = bias+pos_t */
  long long norm_g, norm_b, head_w, head_b;          /* wf32 */
  ribca_block_desc blocks[16];
} ribca_vit_desc;

size_t ribca_vit_workspace_bytes(const ribca_vit_desc* desc, int n_cells);
/* patches [n_cells][in_chans][40][40] fp32 -> probs [n_cells][classes] (softmax) and, optionally, logits */
int ribca_vit_forward(const ribca_vit_desc* desc, const float* wf32, const void* wsplit,
                      const float* patches, int n_cells, float* probs, float* logits,
                      void* workspace, size_t workspace_bytes, int precision, ribca_stream_t stream);

typedef struct ribca_mae_desc {
  int channels;                    /* L = grid rows * cols (7 / 10 / 15) */
  int enc_dim, enc_heads, enc_depth, dec_dim, dec_heads, dec_depth;
  int plane_format, w_log2_scale;  /* as in ribca_vit_desc */
  long long split_plane;
  long long embed_w;               /* wsplit [enc_dim][1600] */
  long long embed_bias, cls_token, pos_embed;            /* wf32: [enc_dim], [enc_dim], [1+L][enc_dim] */
  long long norm_g, norm_b;
  long long dec_embed_w;           /* wsplit [dec_dim][enc_dim] */
  long long dec_embed_b, mask_token, dec_pos_embed;      /* wf32 */
  long long dec_norm_g, dec_norm_b;
  long long pred_w;                /* wsplit [1600][dec_dim] */
  long long pred_b;
  ribca_block_desc enc_blocks[16];
  ribca_block_desc dec_blocks[16];
} ribca_mae_desc;

size_t ribca_mae_workspace_bytes(const ribca_mae_desc* desc, int n_cells);
/* in-place imputation of the missing channels of patches [n_cells][L][40][40];
 * h_present[n_present] = ascending positions of the markers that are present. */
int ribca_mae_impute(const ribca_mae_desc* desc, const float* wf32, const void* wsplit,
                     float* patches, int n_cells, const int* h_present, int n_present,
                     void* workspace, size_t workspace_bytes, int precision, ribca_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 5 - vote merge, "Others" threshold, per-type counts.
 *        replaces Annotator.merge_by_voting + get_void_vote  cta/model.py:481-636, cta/utils.py:143-146
 *
 * Up to two probability tables take part (the reference's elif chain never merges three).
 * Class k of model m is global cell type h_type_of_class[m][k] (index into the 18-entry list of
 * cta/model.py:97-99, 17 = "Others").  h_vote_rank[type] = position in get_void_vote()'s key order
 * (tie-break: first maximum in that order for two models; class-index order for one model).
 * h_type_thresh[18] = cell_type_confidence values; confidence = global threshold.
 * Outputs: label[n] (uint8 global type), conf[n] (float32, -1 when re-labelled "Others"),
 * counts[18] (int64, accumulated: caller zeroes; may be null), margin[n] (float32, may be null): the decision
 * margin of the cell = the smallest change of its probabilities that could change (label, re-labelled?) -
 * min(|winner - its threshold|, gap to every candidate that would give another outcome).  Cells whose margin
 * is below the error bound of a reduced-precision forward are re-evaluated at higher precision by the host
 * (pipeline.refine_labels), which is what makes the labels of cta/model.py:481-636 exact by construction.
 */
int ribca_merge_votes(const float* probs0, int classes0, const int* h_type_of_class0,
                      const float* probs1, int classes1, const int* h_type_of_class1, int n_cells,
                      const int* h_vote_rank, const float* h_type_thresh, float confidence,
                      uint8_t* label, float* conf, long long* counts, float* margin, ribca_stream_t stream);

/* Per-pixel map from per-cell values (label / colour / confidence maps of Annotator.colorize,
 * cta/model.py:806-858): out[p*channels + c] = mask[p] > 0 ? cell_value[id_to_index[mask[p]]*channels + c] : 0. */
int ribca_paint_cells(const int32_t* mask, long long n_pixels, const int32_t* id_to_index, int max_id,
                      const uint8_t* cell_value, int channels, uint8_t* out, ribca_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Spatial statistics on the labelled cells.
 *   ribca_knn_2d: exact k nearest neighbours (self included, as sklearn's kneighbors on the training set) of n
 *     points in the plane.  replaces NearestNeighbors(..., algorithm='ball_tree').kneighbors
 *     (cta/spatial_methods.py:35-37, 97-99, 153-155).  The caller bins the points into a uniform gx x gy grid of edge
 *     `cell` anchored at (x0, y0): xy_sorted[n][2] (float64) are the points in ascending bin order (bin = by*gx+bx),
 *     order[s] = original index of sorted point s, bin_start[gx*gy+1] the bin offsets.  out_idx[n][k] (original
 *     indices, row = original query index) ascending by (float64 distance, index); out_d2 (optional) squared distances.
 *   ribca_neighbor_stats: from nbr[n][k] and types[n] (0..n_types-1), ignoring the first `skip` neighbours (the
 *     point itself):  type_matrix[a][b] += #{(j, m): type[j]=a, type[nbr[j][m]]=b}  (spatial_methods.py:36-40; caller
 *     zeroes / accumulates over images), and compositions[n][n_levels*n_types] = per level L the type histogram of
 *     the L nearest neighbours divided by its sum (spatial_methods.py:157-176).  Either output may be null.
 */
int ribca_knn_2d(const double* xy_sorted, const int* order, const int* bin_start, int n, int k,
                 double x0, double y0, double cell, int gx, int gy, int* out_idx, double* out_d2,
                 ribca_stream_t stream);
int ribca_neighbor_stats(const int* nbr, const int* types, int n, int k, int skip, int n_types,
                         unsigned long long* type_matrix, const int* h_levels, int n_levels,
                         double* compositions, ribca_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RIBCA_B200_H */
