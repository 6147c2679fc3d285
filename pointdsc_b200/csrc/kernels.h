// Internal launch API of the pointdsc_b200 kernels.  Host-callable; every function only enqueues
// work on `st`.  Shapes: B sets, N correspondences per set, NS = SC row stride (N rounded up to 64),
// C = 128 channels, S seeds per set, k neighbours per seed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdsc {

// ---- stage i ------------------------------------------------------------------------------------
void launch_sc_matrix(const float* src, const float* tgt, float* sc, int B, int N, int NS, float sigma_d,
                      cudaStream_t st);

// tensor-core path: sc_t[b][kt][qt][16][128][4] tiles (see sc_matrix.cu); size B * ceil(N/64) * ceil(N/128) * 8192 floats
void launch_sc_matrix_tiled(const float* src, const float* tgt, float* sc, int B, int N, float sigma_d, cudaStream_t st);
void launch_sc_untile(const float* sc_t, float* out, int B, int N, cudaStream_t st);

// ---- stage ii, fp32 SIMT path -------------------------------------------------------------------
// out[b][r][o] = epi( sum_c A[b][r][c] * W[b][o][c] )   A,W K-contiguous; K % 16 == 0.
//   epi 0: (+bias[o]) (relu) (+res[r][o])     epi 1: 2 - 2*acc  (feature-space distance, common.py:58-61)
//   epi 2: clamp(1 - (1 - acc) / epi_param, 0, 1) with a zero diagonal  (feature similarity M, PointDSC.py:160-165)
struct LinearArgs {
  const float* A; long long strideA; int lda;
  const float* W; long long strideW; int ldw;
  const float* bias; const float* res; int ldres;
  float* out; long long strideO; int ldo;
  int M, K, Nout, relu, epi, batch;
  float epi_param;
};
void launch_linear_simt(const LinearArgs& a, cudaStream_t st);
void launch_layer0(const float* corr_pos, const float* W, const float* bias, float* out, long long rows, int in_dim,
                   cudaStream_t st);
void launch_attention_simt(const float* q, const float* k, const float* v, const float* sc, float* msg, int B, int N,
                           int NS, cudaStream_t st);

// ---- a4 + a5: normalise + classification head ---------------------------------------------------
struct HeadWeights {
  const float* w0t;  // [128][32]  classification.0.weight transposed
  const float* b0;   // [32]
  const float* w2t;  // [32][32]   classification.2.weight transposed
  const float* b2;   // [32]
  const float* w4;   // [32]
  const float* b4;   // [1]
};
void launch_head(const float* feat, const HeadWeights& w, float* normed, float* conf, long long rows, int want_conf,
                 cudaStream_t st);

// ---- a6: seeds -----------------------------------------------------------------------------------
void launch_pick_seeds(const float* src, const float* conf, int32_t* seeds, float* key_scratch, int B, int N, int S,
                       float radius, cudaStream_t st);
void launch_top_seeds(const float* conf, int32_t* seeds, int B, int N, int S, cudaStream_t st);   // a6' (non-testing rule)
int pick_seeds_max_n();

// ---- a7: seed-row kNN ----------------------------------------------------------------------------
void launch_gather_rows(const float* normed, const int32_t* seeds, float* out, int B, int N, int S, cudaStream_t st);
// tensor-core seed-row distances (knn_tc.cu): dist[b][s][j] = 2 - 2 <normed[b][seeds[b][s]], normed[b][j]>, fp16 hi/lo split
void launch_knn_dist_tc(const float* normed, const int32_t* seeds, float* dist, int B, int N, int S, cudaStream_t st);
void launch_knn_select(const float* dist, int32_t* knn_idx, int B, int N, int S, int k, cudaStream_t st);

// ---- a8 + a9: compatibility + power iteration -----------------------------------------------------
void launch_nsm_power(const float* normed, const float* src, const float* tgt, const int32_t* knn_idx, float* iterates,
                      uint32_t* conv_mask, float* compat_out, int B, int N, int S, int k, int iters, float sigma,
                      float sigma_d, int mask_stride, int tensor_gram, cudaStream_t st);   // tensor_gram: fp16 hi/lo mma.sync Gram (k <= 40)

// ---- a10 + a11: weighted Kabsch per seed, hypothesis scoring, selection -----------------------------
void launch_seed_hypotheses(const float* src, const float* tgt, const int32_t* knn_idx, const float* iterates,
                            const uint32_t* conv_mask, const float* seed_trans_in, float* seed_trans,
                            int32_t* inlier_counts, unsigned long long* best_key, float* eig_out, int32_t* power_iters,
                            int B, int N, int S, int k, int iters, float inlier_threshold, int mask_stride,
                            cudaStream_t st);

// ---- a11 (labels) + a12: refinement ----------------------------------------------------------------
void launch_select_refine(const float* src, const float* tgt, const float* seed_trans,
                          const unsigned long long* best_key, float* final_trans, float* final_labels,
                          float* init_trans_out, int32_t* best_out, int32_t* refine_solves, int B, int N, int S,
                          float inlier_threshold, float refine_threshold, int max_refine, cudaStream_t st);

// ---- f3: per-pair evaluation statistics (eval_stats.cu), 10 floats per set -----------------------------------
void launch_eval_stats(const float* pred_trans, const float* gt_trans, const float* src, const float* tgt,
                       const float* pred_labels, const float* gt_labels, float* stats, int B, int N, float re_thre,
                       float te_thre, cudaStream_t st);

// ---- f1: correspondence front end (frontend.cu): nearest neighbour in descriptor space, mutual check, centred input ----
size_t match_scratch_bytes(int Ns, int Nt);
int match_max_dim();
void launch_match(const void* src_desc, const void* tgt_desc, int desc_is_fp64, const float* src_keypts, const float* tgt_keypts,
                  int Ns, int Nt, int D, int mutual, void* scratch, int32_t* corr, int32_t* count, float* corr_pos,
                  float* out_src, float* out_tgt, cudaStream_t st);

// ---- f4: N x N power iteration (eig_power.cu) -----------------------------------------------------------------
size_t eig_scratch_bytes(int B, int N);
int launch_leading_eigenvector(const float* M, float* v, int* iters_run, int B, int N, int iters, int early_exit, void* scratch,
                               cudaStream_t st);   // returns cudaError_t

// ---- f2: descriptor front end (fpfh.cu): voxel down-sampling, normals, FPFH -------------------------------------
size_t voxel_scratch_bytes(long long n);
void launch_voxel_down_sample(const float* pts, long long n, double voxel, float* out_pts, int32_t* out_count, int32_t* status,
                              void* scratch, cudaStream_t st);
size_t fpfh_scratch_bytes(int m, int max_nn);
int launch_estimate_normals(const float* pts, int m, double radius, int max_nn, double* normals, int32_t* status, void* scratch,
                            cudaStream_t st);      // returns cudaError_t
int launch_compute_fpfh(const float* pts, const double* normals, int m, double radius, int max_nn, int normalise, double* out,
                        int32_t* status, void* scratch, cudaStream_t st);      // returns cudaError_t

// ---- per-device launch configuration (device_state.cu) ----------------------------------------------------
// opt `kernel` in to `bytes` of dynamic shared memory on the CURRENT device (no-op if already granted there)
cudaError_t ensure_dynamic_smem(const void* kernel, int bytes);
int device_sm_count();   // SM count of the current device

// ---- misc ---------------------------------------------------------------------------------------
void launch_fill_u32(uint32_t* p, uint32_t v, long long n, cudaStream_t st);
void launch_fill_u64(unsigned long long* p, unsigned long long v, long long n, cudaStream_t st);

}  // namespace pdsc
