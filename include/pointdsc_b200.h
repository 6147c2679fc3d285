/*
 * pointdsc_b200 — C ABI of the B200-native PointDSC testing-mode forward engine.
 *
 * This is the drop-in boundary for ONE path of the reference: `PointDSC.forward` with the
 * 'testing' key set (reference models/PointDSC.py:128-197).  The reference has no FFI of its
 * own (it is pure Python/PyTorch); the binding a maintainer adds is the ctypes stub in
 * pointdsc_b200/_capi.py (shown in INTEGRATION.md), driven by a torch.nn.Module with the
 * reference's constructor, forward(dict)->dict and state_dict keys (pointdsc_b200/model.py).
 *
 * Conventions
 *   - plain C types only; every entry point returns 0 on success or a pdsc_status code and
 *     records a message retrievable with pdsc_last_error() (thread-local).
 *   - "d_" pointers are device pointers on the engine's device, "h_" pointers host pointers.
 *   - all device work is enqueued on the caller's stream; pdsc_forward() performs NO host
 *     synchronisation and no allocation, so it can be captured in a CUDA graph.
 *   - tensors are dense, row-major, fp32 unless stated; B = number of correspondence sets
 *     in the call, N = correspondences per set (the same N for the whole call).
 *   - a batched call is, by definition, the loop of per-set testing forwards (the reference
 *     asserts bs == 1 in testing mode, PointDSC.py:210, :414); in particular the power
 *     iteration's early exit is decided per set (PointDSC.py:354).
 */
#ifndef POINTDSC_B200_H_
#define POINTDSC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdsc_engine pdsc_engine;

typedef enum pdsc_status {
  PDSC_OK = 0,
  PDSC_ERR_INVALID_ARGUMENT = 1,
  PDSC_ERR_UNKNOWN_PARAM = 2,
  PDSC_ERR_SHAPE = 3,
  PDSC_ERR_NOT_COMMITTED = 4,
  PDSC_ERR_WORKSPACE = 5,
  PDSC_ERR_CUDA = 6,
  PDSC_ERR_UNSUPPORTED = 7
} pdsc_status;

/* Arithmetic of the encoder's contractions (stage ii).  All other stages are fp32. */
typedef enum pdsc_precision {
  PDSC_FP32_SIMT = 0, /* fp32 FFMA kernels: the exact-arithmetic path                              */
  PDSC_BF16X3 = 1,    /* tcgen05 kind::f16, bf16 hi/lo operand split, 3 products, fp32 accumulate:
                         16 significant bits per operand                                             */
  PDSC_BF16 = 2,      /* tcgen05 kind::f16, single bf16 operands, fp32 accumulate: throughput mode;
                         the 12-layer near-argmax attention amplifies bf16 rounding, so R/t may
                         deviate from the reference by > 1e-4 on some sets           */
  PDSC_FP16X3 = 3     /* tcgen05 kind::f16, fp16 hi/lo operand split, 3 products, fp32 accumulate:
                         22 significant bits per operand — fp32-grade results on the tensor cores   */
} pdsc_precision;

/* Mirrors PointDSC.__init__ (reference models/PointDSC.py:81-91) plus the engine's precision. */
typedef struct pdsc_config {
  int32_t in_dim;            /* 6                                                        */
  int32_t num_layers;        /* 12 in the released snapshots (ctor default 6)            */
  int32_t num_channels;      /* 128 (only value supported by the kernels)                */
  int32_t num_iterations;    /* power-iteration cap, 10                                  */
  float ratio;               /* seeds = int(N * ratio), 0.1                              */
  float inlier_threshold;    /* hypothesis scoring threshold; also selects the refinement
                                threshold: 0.10 iff == 0.10f else 1.2 (PointDSC.py:415)   */
  float sigma_d;             /* initial value of the `sigma_spat` buffer; a loaded
                                state dict overrides it, as in the reference              */
  int32_t k;                 /* neighbourhood size of the NSM module, 40                 */
  float nms_radius;          /* seed NMS radius                                          */
  int32_t precision;         /* pdsc_precision                                           */
  int32_t device;            /* CUDA device ordinal                                      */
} pdsc_config;

/* Optional taps and injection points at the stage boundaries of SURVEY.md §8(a).  Every pointer may
 * be NULL.  `in_*` tensors REPLACE the engine's own result of that stage (used by the parity tests
 * to feed a stage the reference's upstream tensors); `out_*` tensors receive a copy. */
typedef struct pdsc_stage_io {
  /* injection */
  const float* in_features;     /* [B,N,C]  un-normalised encoder output (skips stages i-ii)         */
  const float* in_confidence;   /* [B,N]    confidence logits (requires in_features)                 */
  const int32_t* in_seeds;      /* [B,S]    seed indices                                            */
  const int32_t* in_knn_idx;    /* [B,S,k]  neighbourhoods                                          */
  const float* in_seed_trans;   /* [B,S,4,4] hypotheses                                             */
  /* taps */
  float* out_sc;                /* [B,N,N]  spatial-consistency matrix (a1)                          */
  float* out_features;          /* [B,N,C]  encoder output (a2-a3)                                   */
  float* out_normed;            /* [B,N,C]  L2-normalised features (a4)                              */
  float* out_confidence;        /* [B,N]    (a5)                                                     */
  int32_t* out_seeds;           /* [B,S]    (a6)                                                     */
  int32_t* out_knn_idx;         /* [B,S,k]  (a7)                                                     */
  float* out_compat;            /* [B,S,k,k] (a8)                                                    */
  float* out_eig;               /* [B,S,k]  leading eigenvector at the set's exit iteration (a9)     */
  int32_t* out_power_iters;     /* [B]      iterations run per set (a9)                              */
  float* out_seed_trans;        /* [B,S,4,4] (a10)                                                   */
  int32_t* out_inlier_counts;   /* [B,S]    inlier count of every hypothesis (a11)                   */
  int32_t* out_best;            /* [B]      selected hypothesis (a11)                                */
  float* out_init_trans;        /* [B,4,4]  selected hypothesis before refinement (a11)              */
  int32_t* out_refine_solves;   /* [B]      Kabsch solves done by the refinement (a12)               */
  int32_t layer_tap;            /* if out_layer_features != NULL: which encoder layer to copy        */
  float* out_layer_features;    /* [B,N,C]  output of encoder layer `layer_tap`                      */
  float* out_layer_debug;       /* [5,B,N,C] internals of layer `layer_tap`: PointCN output, q, k, v, msg.
                                   In the tensor-core modes q carries the folded log2(e)/sqrt(C) scale. */
  int64_t* out_timeline;        /* [2,16,4,8] clock64() stamps of CTA 0 of the layer-`layer_tap` PointCN+Q chain
                                   kernel and attention kernel (tensor-core modes; developer tool)            */
} pdsc_stage_io;

/* ---- lifetime --------------------------------------------------------------------------------- */
int pdsc_create(const pdsc_config* cfg, pdsc_engine** out);
int pdsc_destroy(pdsc_engine* e);
const char* pdsc_last_error(void);
const char* pdsc_version(void);

/* ---- parameters: the reference's state dict, key for key (PointDSC.py:93-113) ------------------
 * `name` is a state-dict key ("encoder.layer0.weight", "sigma_spat", ...); `h_data` holds `count`
 * fp32 values in the tensor's own row-major order.  Keys the path does not use
 * (num_batches_tracked, the stray `gamma`) are accepted and ignored, so a released snapshot can be
 * pushed unfiltered.  pdsc_commit_params() folds eval-mode BatchNorm into the preceding 1x1 conv,
 * builds the device-side operand images and must be called before pdsc_forward(). */
int pdsc_set_param(pdsc_engine* e, const char* name, const float* h_data, int64_t count);
int pdsc_commit_params(pdsc_engine* e);
int pdsc_set_precision(pdsc_engine* e, int32_t precision);

/* ---- sizes ------------------------------------------------------------------------------------- */
int32_t pdsc_num_seeds(const pdsc_engine* e, int32_t N);       /* S = int(N * ratio)          */
int32_t pdsc_num_neighbours(const pdsc_engine* e, int32_t N);  /* k = min(cfg.k, N - 1)       */
size_t pdsc_workspace_bytes(const pdsc_engine* e, int32_t B, int32_t N);

/* ---- the path: PointDSC.forward, testing mode (PointDSC.py:128-197) ---------------------------
 * d_corr_pos [B,N,6], d_src_keypts [B,N,3], d_tgt_keypts [B,N,3]  ->
 * d_final_trans [B,4,4] (maps src onto tgt), d_final_labels [B,N] in {0,1}.
 * `io` may be NULL.  `d_workspace` must hold pdsc_workspace_bytes(e,B,N) bytes, 256-byte aligned,
 * and is only used for the duration of the enqueued work. */
int pdsc_forward(pdsc_engine* e, int32_t B, int32_t N, const float* d_corr_pos, const float* d_src_keypts,
                 const float* d_tgt_keypts, float* d_final_trans, float* d_final_labels,
                 const pdsc_stage_io* io, void* d_workspace, size_t workspace_bytes, void* cuda_stream);

/* pdsc_forward as ONE graph launch: the first call with a given (B, N, buffer addresses) runs eagerly and captures the
 * forward's kernels into a CUDA graph; later calls with the same arguments replay it (one cudaGraphLaunch instead of ~60
 * kernel launches: the small-batch / bs = 1 case of the evaluation loops, evaluation/test_3DMatch.py:133).  The caller keeps
 * the buffers alive and at the same addresses; up to 8 graphs are cached per engine.  Inside a foreign stream capture, or
 * with profiling enabled, it degrades to pdsc_forward. */
int pdsc_forward_graph(pdsc_engine* e, int32_t B, int32_t N, const float* d_corr_pos, const float* d_src_keypts,
                       const float* d_tgt_keypts, float* d_final_trans, float* d_final_labels, void* d_workspace,
                       size_t workspace_bytes, void* cuda_stream);

/* Same call with HOST buffers (the end-to-end form): copies the inputs host->device, runs
 * pdsc_forward, copies the two outputs device->host and synchronises the stream before returning.
 * Device staging and workspace are owned by the engine and grown on demand.  The key points are copied
 * on `cuda_stream`; corr_pos (half of the input bytes) is copied on an engine-owned side stream, ordered
 * behind `cuda_stream` by an event, while the spatial-consistency kernel (which reads only the key points)
 * runs, and joined again before the first kernel that reads corr_pos. */
int pdsc_forward_host(pdsc_engine* e, int32_t B, int32_t N, const float* h_corr_pos, const float* h_src_keypts,
                      const float* h_tgt_keypts, float* h_final_trans, float* h_final_labels, void* cuda_stream);

/* The host form as a two-deep pipeline, for loops over many batches (the evaluation drivers' `for data in loader`,
 * evaluation/test_3DMatch.py:64-101, whose DataLoader workers prefetch the next batch while the model runs the current one):
 * _submit enqueues the host->device copies of this call's inputs on an engine-owned copy stream and the forward on
 * `cuda_stream` behind them, then returns WITHOUT synchronising; `*slot_out` (0 or 1) names the call.  _wait(slot) blocks until
 * that call's forward has finished, copies the two results into the host buffers given to _submit (on a second engine-owned
 * stream, i.e. beside the next call's forward) and returns when they have arrived.  Two calls may be in flight: the input copies
 * of call t + 1 and the result copies of call t - 1 then run beside the forward of call t (the forwards themselves stay
 * serialised on `cuda_stream`: they share one workspace).  All five host buffers must stay valid until _wait returns; the INPUT
 * buffers should be page-locked (from a pageable buffer the copy is synchronous: correct, but without the overlap), the result
 * buffers may be pageable.  A third _submit before a _wait fails with PDSC_ERR_INVALID_ARGUMENT.  Results are bit-identical to
 * pdsc_forward_host's. */
int pdsc_forward_host_submit(pdsc_engine* e, int32_t B, int32_t N, const float* h_corr_pos, const float* h_src_keypts,
                             const float* h_tgt_keypts, float* h_final_trans, float* h_final_labels, void* cuda_stream,
                             int32_t* slot_out);
int pdsc_forward_host_wait(pdsc_engine* e, int32_t slot);

/* ---- the same module call WITHOUT the 'testing' key (validation during training, PointDSC.py:158-165, :176, :190-191):
 * seeds are the top-S correspondences by confidence (no suppression), the power iteration's early exit is decided over the
 * whole batch (the reference's allclose spans [bs*S, k]), there is no post-refinement, `d_confidence` [B,N] receives the
 * classification logits (the reference returns them as final_labels) and, if `d_M` is not NULL, it receives the feature
 * similarity matrix M = clamp(1 - (1 - F F^T) / sigma^2, 0, 1) with a zero diagonal, [B,N,N].  Eval-mode BatchNorm only
 * (running statistics): the training-mode forward and the backward pass are outside this engine. */
int pdsc_forward_eval(pdsc_engine* e, int32_t B, int32_t N, const float* d_corr_pos, const float* d_src_keypts,
                      const float* d_tgt_keypts, float* d_final_trans, float* d_confidence, float* d_M,
                      const pdsc_stage_io* io, void* d_workspace, size_t workspace_bytes, void* cuda_stream);

/* ---- next rows of the path ---------------------------------------------------------------
 * f3: per-pair evaluation statistics, replacing libs/loss.py:34-63 (TransformationLoss) + :94-100 (ClassificationLoss,
 * scikit-learn on the host) and the per-pair host synchronisation of evaluation/test_3DMatch.py:83-101.
 * d_stats [B,10] = [success, RE deg, TE cm, #gt inliers, gt inlier ratio, #gt inliers among the kept, precision, recall,
 * f1, rmse]; thresholds as the drivers pass them (3DMatch: 15 deg / 30 cm, KITTI: 5 deg / 60 cm). */
int pdsc_eval_stats(pdsc_engine* e, int32_t B, int32_t N, const float* d_pred_trans, const float* d_gt_trans,
                    const float* d_src_keypts, const float* d_tgt_keypts, const float* d_pred_labels,
                    const float* d_gt_labels, float re_thre, float te_thre, float* d_stats, void* cuda_stream);

/* f1: the correspondence front end, replacing datasets/ThreeDMatch.py:283-291 + :299-308 (the same lines in
 * datasets/KITTI.py:80-114, demo_registration.py:101-108): nearest neighbour of every source descriptor among the target
 * descriptors under sqrt(2 - 2 <a,b> + 1e-6) evaluated in the descriptors' own dtype (desc_is_fp64: 0 = fp32 FCGF,
 * 1 = fp64 FPFH), first minimum wins, optional mutual check, then the in_dim = 6 network input.  One pair per call.
 * Outputs in the layout pdsc_forward consumes, sized for the worst case M = Ns: d_corr [Ns,2] int32 (source, target),
 * d_count [1] = M, d_corr_pos [Ns,6] (centred), d_out_src / d_out_tgt [Ns,3]; only the first M rows are written.
 * D <= 64; d_scratch holds pdsc_match_scratch_bytes(Ns, Nt) bytes, 8-byte aligned. */
size_t pdsc_match_scratch_bytes(int32_t Ns, int32_t Nt);
int pdsc_match(pdsc_engine* e, int32_t Ns, int32_t Nt, int32_t D, const void* d_src_desc, const void* d_tgt_desc,
               int32_t desc_is_fp64, const float* d_src_keypts, const float* d_tgt_keypts, int32_t use_mutual,
               int32_t* d_corr, int32_t* d_count, float* d_corr_pos, float* d_out_src, float* d_out_tgt, void* d_scratch,
               size_t scratch_bytes, void* cuda_stream);

/* f2: the descriptor front end, replacing the open3d 0.9 calls of misc/cal_fpfh.py:21-26 (and demo_registration.py:37-44):
 *   pcd.voxel_down_sample(voxel)                                                -> pdsc_voxel_down_sample
 *   pcd.estimate_normals(KDTreeSearchParamHybrid(radius = 2 voxel, max_nn = 30)) -> pdsc_estimate_normals
 *   compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(5 voxel, 100))             -> pdsc_compute_fpfh
 *   o3d.io.read_point_cloud(path).points                                         -> pdsc_read_ply (host)
 * open3d is not part of the reference tree: these follow its published algorithms (oracle/fpfh_oracle.py; PARITY UNPINNED).
 * d_points [n,3] float32.  pdsc_voxel_down_sample writes the voxel means to d_out_points (room for [n,3]; rows in ascending
 * (ix, iy, iz) voxel order) and their number to d_count[0]; read d_count after synchronising the stream.  Normals are float64
 * [m,3] (largest-magnitude component positive; (0,0,1) below three neighbours), FPFH float64 [m,33] — the dtype pdsc_match
 * takes with desc_is_fp64 = 1 — with normalise != 0 applying x / (||x|| + 1e-6) per row (demo_registration.py:43).
 * d_status[0] is a bit mask written on the stream: 1 = more than 2^21 voxels along an axis or a non-finite coordinate,
 * 2 = a neighbourhood holds more than 4096 points inside the radius (the search is brute force over the m key points and
 * sized for down-sampled clouds).  Scratch: 8-byte aligned, *_scratch_bytes() bytes. */
size_t pdsc_voxel_down_sample_scratch_bytes(int64_t n);
int pdsc_voxel_down_sample(pdsc_engine* e, int64_t n, const float* d_points, double voxel_size, float* d_out_points,
                           int32_t* d_count, int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* cuda_stream);
size_t pdsc_fpfh_scratch_bytes(int32_t m, int32_t max_nn);
int pdsc_estimate_normals(pdsc_engine* e, int32_t m, const float* d_points, double radius, int32_t max_nn, double* d_normals,
                          int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* cuda_stream);
int pdsc_compute_fpfh(pdsc_engine* e, int32_t m, const float* d_points, const double* d_normals, double radius, int32_t max_nn,
                      int32_t normalise, double* d_fpfh, int32_t* d_status, void* d_scratch, size_t scratch_bytes, void* cuda_stream);
/* Vertex positions of a PLY file (ascii or binary_little_endian; x, y, z float or double) into host memory as [n,3] float32.
 * Call with points = NULL to learn *n_vertices, then with a buffer of `capacity` >= n vertices. */
int pdsc_read_ply(const char* path, float* points, int64_t capacity, int64_t* n_vertices);

/* f4: leading eigenvector of N x N compatibility matrices by power iteration, replacing cal_leading_eigenvector(M, 'power')
 * (models/PointDSC.py:338-358) in its N x N uses: the learned feature-similarity matrix of the non-testing forward (:170) and
 * the classical spectral-matching baseline (baseline_scripts/baseline_3DMatch.py:40-44, ten fixed iterations = early_exit 0).
 * d_M [B,N,N] row-major; start vector all ones; v <- M v / (||M v|| + 1e-6); with early_exit the iteration of a set stops
 * when allclose(v, v_prev, rtol 1e-5, atol 1e-8) holds (per set, decided on the device).  d_eigenvector [B,N],
 * d_iterations_run [B]; d_scratch holds pdsc_leading_eigenvector_scratch_bytes(B, N) bytes, 16-byte aligned. */
size_t pdsc_leading_eigenvector_scratch_bytes(int32_t B, int32_t N);
int pdsc_leading_eigenvector(pdsc_engine* e, int32_t B, int32_t N, const float* d_M, int32_t num_iterations, int32_t early_exit,
                             float* d_eigenvector, int32_t* d_iterations_run, void* d_scratch, size_t scratch_bytes,
                             void* cuda_stream);

/* ---- live profiling with CUDA events on the caller's stream ------------------------------------------
 * When enabled, pdsc_forward() records an event pair around each stage below (and around EVERY launch of
 * the dominant kernel, the per-layer attention).  pdsc_profile_read() waits for the last forward's events
 * and returns, per span, the accumulated milliseconds and the number of launches since the last read.
 * Not capturable in a CUDA graph; costs two event records per span. */
typedef enum pdsc_span {
  PDSC_SPAN_SC = 0,          /* a1  sc_matrix                                   */
  PDSC_SPAN_LINEAR = 1,      /* a2  layer0 + PointCN/QKV/fc_message kernels     */
  PDSC_SPAN_ATTENTION = 2,   /* a3  attention kernel (one span per layer)       */
  PDSC_SPAN_HEAD = 3,        /* a4+a5                                           */
  PDSC_SPAN_SEEDS = 4,       /* a6                                              */
  PDSC_SPAN_KNN = 5,         /* a7                                              */
  PDSC_SPAN_NSM = 6,         /* a8+a9                                           */
  PDSC_SPAN_HYPOTHESES = 7,  /* a10+a11                                         */
  PDSC_SPAN_REFINE = 8,      /* a11 labels + a12                                */
  PDSC_SPAN_TOTAL = 9,       /* whole pdsc_forward                              */
  PDSC_SPAN_COUNT = 10
} pdsc_span;
int pdsc_profile_enable(pdsc_engine* e, int32_t enable);
int pdsc_profile_read(pdsc_engine* e, float* ms_out /* [PDSC_SPAN_COUNT] */, int32_t* launches_out /* [PDSC_SPAN_COUNT] */);

/* Number of kernels one pdsc_forward(B,N) call launches at the current precision (bench.py's
 * `gpu_launches`). */
int32_t pdsc_launches_per_forward(const pdsc_engine* e, int32_t B, int32_t N);

#ifdef __cplusplus
}
#endif
#endif /* POINTDSC_B200_H_ */
