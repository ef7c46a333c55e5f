/* ape_b200.h -- C-ABI of the B200-native 6D-pose geometry hot path.
 *
 * Drop-in boundary for KochPJ/AutoPoseEstimation main.py options 4 ("Create Pose
 * labels") and 6 ("Run Live Prediction").  The reference has one native entry point
 * (`int knn(at::Tensor&, at::Tensor&, at::Tensor&)`, DenseFusion/lib/knn/src/knn.h:12,
 * exported by src/vision.cpp:3-5); everything else on the path is Python that calls
 * torch / numpy / open3d.  Each function below names the reference interface it
 * replaces.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - calls are asynchronous on `stream`; the caller owns all buffers;
 *   - return value: APE_OK or an APE_ERR_* code; ape_last_error() gives the message of
 *     the last failing call on the calling thread.  Nothing aborts the process
 *     (the reference's knn.h:41-46 printf + THError abort is replaced by a status).
 *   - there is NO CPU fallback: without a CUDA device every compute call returns
 *     APE_ERR_CUDA.
 */
#ifndef APE_B200_H
#define APE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    APE_OK = 0,
    APE_ERR_INVALID = 1,      /* bad argument (null pointer, non-positive size, unsupported shape) */
    APE_ERR_CUDA = 2,         /* CUDA runtime error (message in ape_last_error) */
    APE_ERR_CAPACITY = 3,     /* an output/scratch capacity supplied by the caller is too small */
    APE_ERR_UNSUPPORTED = 4   /* valid request outside what the kernels implement */
};

/* kNN distance arithmetic (both are fp32, summed in dimension order):
 *   APE_KNN_ARITH_CPU  d += (r-q)*(r-q) with product and sum rounded separately
 *                      = DenseFusion/lib/knn/src/cpu/knn_cpu.cpp:8-16 as gcc compiles it
 *   APE_KNN_ARITH_FMA  d = fma(r-q, r-q, d)
 *                      = DenseFusion/lib/knn/src/cuda/knn.cu:86-92 as nvcc compiles it   */
enum { APE_KNN_ARITH_CPU = 0, APE_KNN_ARITH_FMA = 1 };

int         ape_version(void);
const char* ape_last_error(void);
/* Number of kernel launches issued through this library by the calling process so far. */
uint64_t    ape_launch_count(void);
/* Optional per-launch device timing (used by bench.py for the roofline): when enabled, every kernel
 * launch is bracketed by CUDA events on its stream.  ape_profile_report writes one line per kernel
 * label, "label launches total_ms", into buf and returns the number of bytes needed.           */
int         ape_profile_enable(int on);
int         ape_profile_report(char* buf, int buflen);

/* ---------------------------------------------------------------------------------------------
 * a3. DenseFusion back-projection at fixed sampling indices.
 * Replaces pipeline/utils.py:542-553 (= datasets/myDatasetAugmented/dataset.py:260-275).
 *   depth      [n_frames,height,width] u16 raw sensor units
 *   frame_of   [n_obj] frame index of each object, or NULL (object b reads frame b)
 *   bbox       [n_obj,4] rmin,rmax,cmin,cmax from get_bbox (dataset.py:342-380)
 *   choose     [n_obj,n_points] int64 flat indices into the bbox crop (width cmax-cmin)
 *   cam        [n_obj,5] fp32 ppx,ppy,fx,fy,depth_scale (already rounded to fp32, as numpy does)
 *   cloud      [n_obj,n_points,3] fp32 out: x=((col-ppx)*z)/fx, y=((row-ppy)*z)/fy, z=d*scale,
 *              each operation rounded to fp32 in that order (bit-exact with the reference)  */
int ape_backproject_choose(const uint16_t* depth, int n_frames, int height, int width,
                           const int32_t* frame_of, const int32_t* bbox, const int64_t* choose,
                           const float* cam, int n_obj, int n_points, float* cloud, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a1 + a2 + a3 on the device: per detected object, label mask -> get_bbox
 * (datasets/myDatasetAugmented/dataset.py:342-380) -> choose (pipeline/utils.py:524-539) -> fp32
 * back-projection (:542-553), without the host round trip and without the xmap/ymap lists (:518-519).
 *   label [n_frames,H,W] u8, depth [n_frames,H,W] u16; frame_of [n_obj] or NULL; label_value [n_obj] or NULL (= 255)
 *   seeds [n_obj] u32 or NULL: more than n_points candidates -> the n_points with the smallest keys
 *         mix32(seed ^ c * 0x9E3779B9) (murmur3 finaliser; ties: lower index) are kept, ascending -- the distribution of
 *         the reference's np.random.shuffle subset, bit-exact against oracle/geometry.py; at most n_points -> 'wrap'
 *   cam [n_obj,5] fp32 (ppx,ppy,fx,fy,depth_scale)
 *   bbox [n_obj,4] out (rmin,rmax,cmin,cmax), n_candidates [n_obj] out (0 = object skipped, as :530-531),
 *   choose [n_obj,n_points] int64 out, cloud [n_obj,n_points,3] fp32 out or NULL                          */
int ape_mask_bbox_choose(const uint8_t* label, const uint16_t* depth, int n_frames, int height, int width,
                         const int32_t* frame_of, const uint8_t* label_value, const uint32_t* seeds,
                         const float* cam, int n_obj, int n_points, int32_t* bbox, int32_t* n_candidates,
                         int64_t* choose, float* cloud, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4. Masked depth -> point cloud in the robot frame (ordered stream compaction).
 * Replaces the per-pixel loop of pc_reconstruction/open3d_utils.py:172-192 (get_surface).
 * A "view" is one (frame, label value) pair; pixels are emitted in row-major order
 * (np.where order, :174) when label matches and depth != 0.
 *   label        [n_frames,height,width] u8;  depth [n_frames,height,width] u16 (raw = mm)
 *   frame_of     [n_views] frame index, or NULL (view v reads frame v)
 *   label_value  [n_views] u8: 0 = "label != 0" (the reference's test), k>0 = "label == k"
 *   cam          [n_views,4] fp64 ppx,ppy,fx,fy;  robot2cam [n_views,16] fp64 row-major 4x4
 *   capacity     max points per view;  points [n_views,capacity,3] fp64 out
 *   pixel_index  [n_views,capacity] int32 flat pixel index of every emitted point, or NULL
 *   counts       [n_views] int32 out: number of valid pixels (may exceed capacity: the excess is dropped,
 *                the caller checks)
 *   work         scratch of ape_surface_work_bytes(n_views, height, width) bytes (16-byte aligned; the call
 *                does not need it initialised): per-chunk valid-pixel counts + one validity bit per pixel           */
size_t ape_surface_work_bytes(int n_views, int height, int width);
int ape_surface_backproject(const uint8_t* label, const uint16_t* depth, int n_frames, int height, int width,
                            const int32_t* frame_of, const uint8_t* label_value,
                            const double* cam, const double* robot2cam, int n_views, int capacity,
                            double* points, int32_t* pixel_index, int32_t* counts, void* work, void* stream);

/* a4 for frames that carry SEVERAL object labels (BASELINE config 4: 10 k frames x 5 objects): one pass over a frame's
 * label + depth produces the validity bits of all n_labels label values, so the 921 600 B of a frame are read once per
 * frame, not once per (frame, object).  View v = frame * n_labels + l (label value label_values_host[l], non-zero).
 *   cam [n_frames,4], robot2cam [n_frames,16] fp64: per FRAME
 *   points [total_capacity,3] fp64 out, PACKED: view v occupies points[offsets[v] : offsets[v+1]) in row-major pixel order;
 *   counts [n_views] / offsets [n_views+1] int32 out (device); points past total_capacity are dropped (the caller checks
 *   offsets[n_views] <= total_capacity); pixel_index [total_capacity] int32 out or NULL;
 *   work: ape_surface_work_bytes(n_views, height, width) bytes.  No host synchronisation: offsets feed
 *   ape_voxel_down_sample and ape_icp_p2p_ex directly.                                                              */
int ape_surface_backproject_multi(const uint8_t* label, const uint16_t* depth, int n_frames, int height, int width,
                                  const uint8_t* label_values_host, int n_labels, const double* cam, const double* robot2cam,
                                  int total_capacity, double* points, int32_t* pixel_index, int32_t* counts, int32_t* offsets,
                                  void* work, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a14. Brute-force k nearest neighbours.
 * Replaces `int knn(ref, query, idx)` DenseFusion/lib/knn/src/knn.h:12 (python:
 * KNearestNeighbor.forward, lib/knn/__init__.py:15-23).
 *   ref [B,D,N] fp32, query [B,D,M] fp32, idx [B,k,M] int64 out, 1-based, ascending distance,
 *   lowest index first on exact ties.  No scratch (the reference's N*M distance matrix is
 *   never materialised).  D==3,k==1 runs the tiled shared-memory kernel; other shapes run a
 *   generic kernel (k <= 64).                                                                  */
int ape_knn(const float* ref, const float* query, int64_t* idx, int B, int D, int N, int M, int k,
            int arith, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a13/a15. ADD / ADD-S metric, fused (transform + top-1 NN + gather + mean norm).
 * Replaces DenseFusion/lib/loss_refiner.py:39-49 and tools/eval_linemod.py:118-130.
 *   quat [B,4] wxyz (normalised inside, fp32), trans [B,3]
 *   model_points: instance b reads  model_points + b*model_stride  ([n_model,3] fp32; stride in floats, 0 = shared)
 *   target:       instance b reads  target + b*target_stride      ([n_target,3] fp32)
 *   symmetric [B] u8: 1 -> ADD-S (nearest target point per predicted point, APE_KNN_ARITH_CPU
 *   arithmetic), 0 -> ADD: model point i is paired with target row i (requires n_target >= n_model; an instance that
 *   violates it gets dis = NaN, nothing is read out of bounds)
 *   dis [B] fp32 out; nn_index [B,n_model] int32 out (0-based) or NULL                          */
int ape_add_metric(const float* quat, const float* trans, const float* model_points, int64_t model_stride,
                   int n_model, const float* target, int64_t target_stride, int n_target,
                   const uint8_t* symmetric, int B, float* dis, int32_t* nn_index, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5. Point-to-point ICP, whole registration in one persistent kernel.
 * Replaces o3d.registration.registration_icp(source, target, threshold, I,
 * TransformationEstimationPointToPoint(), ICPConvergenceCriteria(rel_fitness, rel_rmse, max_iter))
 * as called by pc_reconstruction/open3d_utils.py:76-104 (open3d 0.9.0 semantics, fp64).
 * Registration r uses source points [src_offset[r], src_offset[r+1]) of `source` and target
 * points [tgt_offset[r], tgt_offset[r+1]) of `target` (both [*,3] fp64, ragged batches).
 *   init          [n_reg,16] fp64 row-major initial transforms, or NULL (identity)
 *   transform     [n_reg,16] fp64 out (row-major 4x4)
 *   info          [n_reg,4]  fp64 out: fitness, inlier_rmse, iterations, n_correspondences (or NULL)
 *   work          scratch of ape_icp_work_bytes(total_source_points, total_target_points) bytes,
 *                 8-byte aligned (working cloud, cell-sorted target, correspondences)            */
size_t ape_icp_work_bytes(int total_source_points, int total_target_points);
int ape_icp_p2p(const double* source, const int32_t* src_offset, const double* target, const int32_t* tgt_offset,
                int n_reg, int total_source_points, int total_target_points,
                double threshold, double rel_fitness, double rel_rmse, int max_iter,
                const double* init, double* transform, double* info, void* work, void* stream);

/* As ape_icp_p2p; src_count [n_reg] int32 or NULL: registration r uses src_count[r] points starting at src_offset[r]
 * (gapped ragged source, e.g. straight from ape_voxel_down_sample's out_points / out_counts).                         */
int ape_icp_p2p_ex(const double* source, const int32_t* src_offset, const int32_t* src_count, const double* target,
                   const int32_t* tgt_offset, int n_reg, int total_source_points, int total_target_points,
                   double threshold, double rel_fitness, double rel_rmse, int max_iter,
                   const double* init, double* transform, double* info, void* work, void* stream);

/* The sequential register-and-merge loop of the object-cloud reconstruction in one call, no host synchronisation.
 * Replaces the loop body of pc_reconstruction/create_pointcloud.py:286-312 (per further non-empty view: icp_regression =
 * voxel grid on both clouds + point-to-point ICP, open3d_utils.py:63-104; transform; concatenate source-first; voxel grid).
 *   points [offset_host[n_views], 3] fp64 (device): the views' surfaces packed; offset_host [n_views + 1] int32 (HOST: the
 *   surface sizes are known when the surfaces are made); empty views are skipped, the first non-empty one starts the cloud.
 *   out_points [offset_host[n_views], 3] fp64, out_count [1] int32 (device): the reconstructed cloud;
 *   out_status [1] int32 (device): 0, or the voxel grid's negative status when an intermediate cloud exceeded
 *   APE_VOXEL_MAX_POINTS (the caller then runs the loop view by view through ape_voxel_down_sample_large);
 *   work: ape_reconstruct_work_bytes(offset_host[n_views]) bytes, 8-byte aligned.                                      */
size_t ape_reconstruct_work_bytes(int total_points);
int ape_reconstruct_run(const double* points, const int32_t* offset_host, int n_views, double voxel_size, double threshold,
                        double rel_fitness, double rel_rmse, int max_iter, double* out_points, int32_t* out_count,
                        int32_t* out_status, void* work, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5/a6. Voxel-grid down-sampling (open3d 0.9 PointCloud::voxel_down_sample semantics; output
 * sorted by voxel index instead of hash-map order).  Replaces pcd.voxel_down_sample(voxel_size)
 * at pc_reconstruction/open3d_utils.py:21, :198 and create_pointcloud.py:312.
 * Ragged batch as above.  out_points has the same capacity/offsets as the input
 * (cloud c writes at most its input count starting at offset[c]); out_counts [n_clouds].
 * The batched kernel sorts a cloud in shared memory: a cloud of more than APE_VOXEL_MAX_POINTS points is NOT processed
 * and gets out_counts[c] = -1 (the call itself returns APE_OK: the sizes live on the device) -- run it through
 * ape_voxel_down_sample_large; out_counts[c] = -2: more than 65 535 (large path: 8 191) voxels along an axis.       */
#define APE_VOXEL_MAX_POINTS 16384
int ape_voxel_down_sample(const double* points, const int32_t* offset, int n_clouds, double voxel_size,
                          double* out_points, int32_t* out_counts, void* stream);
/* ONE cloud of any size up to 2^25 - 1 points (n_points by value), bit-identical output: chunked bitonic sort + merge-path
 * passes + segmented means in global memory (stream-ordered scratch from cudaMallocAsync).  out_count [1] int32 (device). */
int ape_voxel_down_sample_large(const double* points, int n_points, double voxel_size, double* out_points,
                                int32_t* out_count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8/a9. Confidence arg-max, pose of the best point, cloud in the predicted frame.
 * Replaces my_estimator_prediction + get_new_points (DenseFusion/tools/utils.py:7-18, :43-86).
 *   pred_r [B,N,4], pred_t [B,N,3], pred_c [B,N], cloud [B,N,3] fp32
 *   which_max [B] int32 (lowest index on ties), my_r [B,4] (normalised quaternion wxyz),
 *   my_t [B,3], new_points [B,N,3] (may be NULL), pose [B,7] fp64 = (my_r, my_t) widened (may be NULL) */
int ape_pose_select(const float* pred_r, const float* pred_t, const float* pred_c, const float* cloud,
                    int B, int N, int32_t* which_max, float* my_r, float* my_t, float* new_points,
                    double* pose, void* stream);

/* a11. fp64 pose composition of one refinement step.
 * Replaces my_refined_prediction (DenseFusion/tools/utils.py:20-40 -> transformations.py:1254, :1281).
 *   pose_in [B,7] fp64 (wxyz,t); r2 [B,4] fp32 un-normalised; t2 [B,3] fp32; pose_out [B,7] fp64.
 *   If next_points != NULL: cloud [B,N,3] re-expressed in the composed pose as in
 *   tools/eval_linemod.py:92-97 (fp32 R and T) -> next_points [B,N,3].                           */
int ape_pose_compose(const double* pose_in, const float* r2, const float* t2, int B, double* pose_out,
                     const float* cloud, int N, float* next_points, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a7/a10. DenseFusion PoseNet (geometry side) and PoseRefineNet forward, batched over objects.
 * Replaces PoseNet.forward after the colour encoder (DenseFusion/lib/network.py:98-132) and
 * PoseRefineNet.forward (:187-206).  Weights are bound once per model with ape_net_create from
 * the reference state_dict tensors (fp32, reference shapes); the handle owns split-bf16 copies.
 */
typedef struct ape_net ape_net;       /* opaque */

enum { APE_NET_POSENET = 0, APE_NET_REFINER = 1 };
enum { APE_GEMM_TCGEN05 = 0,      /* persistent tcgen05 kernel, 128x256 tiles (gemm_tc2.cuh): the product path */
       APE_GEMM_SIMT = 1,         /* fp32 SIMT validation kernels (tests only) */
       APE_GEMM_TCGEN05_V1 = 2,   /* first generation: non-persistent 128x128 tiles (A/B runs only) */
       APE_GEMM_TCGEN05_PAIR = 3, /* CTA-pair variant, tcgen05.mma.cta_group::2 (gemm_tc3.cuh); measured 5 % slower than the
                                     product path on this workload (profiles/), kept for A/B runs */
       APE_GEMM_TCGEN05_B2B = 4 };/* product path with PoseNet conv1_{r,t,c} -> conv2_{r,t,c} fused back to back in one kernel
                                     (gemm_tc4.cuh): the [R,1920] intermediate never leaves the SM (0.5 GB less HBM traffic per
                                     step); measured 11 % slower on these two layers, kept selectable */

/* weights_host: array of n_tensors host pointers to fp32 tensors in the canonical order listed in
 * DESIGN.md ("weight order"); num_obj as in the reference constructor; max_batch/max_points size the
 * activation workspace (allocated once, on the current device).                                  */
int ape_net_create(int kind, const float* const* weights_host, int n_tensors, int num_obj,
                   int max_batch, int max_points, ape_net** out);
int ape_net_destroy(ape_net* net);
int ape_net_set_gemm(ape_net* net, int gemm_impl);
/* Split-bf16 products per tensor-core layer, in the order conv2|e_conv2, conv5, conv6, heads1, heads2, heads3 (the refiner
 * uses the first three): bit 0 = A_lo*W_hi, bit 1 = A_hi*W_lo, bit 2 = A_hi*W_hi (always set).  7 = all three (fp32-grade
 * products), 6 = the activation's low half dropped, 5 = the weight's low half dropped, 4 = plain bf16.  The defaults are
 * the product configuration chosen from the per-layer error budget; tests and A/B runs change it here.   */
int ape_net_set_passes(ape_net* net, const int* masks6);
int ape_net_get_passes(const ape_net* net, int* masks6);

/* PoseNet geometry forward for B objects.
 *   out_img [B,32,hw] fp32 colour-encoder output per object crop (hw = crop pixels, same for the batch)
 *   cloud [B,N,3] fp32, choose [B,N] int64 (indices into hw), obj [B] int64 class ids
 *   pred_r [B,N,4], pred_t [B,N,3], pred_c [B,N], emb [B,32,N] fp32 out                          */
int ape_posenet_forward(ape_net* net, const float* out_img, int hw, const float* cloud, const int64_t* choose,
                        const int64_t* obj, int B, int N, float* pred_r, float* pred_t, float* pred_c,
                        float* emb, void* stream);

/* PoseRefineNet forward: new_points [B,N,3], emb [B,32,N], obj [B] -> r2 [B,4], t2 [B,3] fp32 */
int ape_refiner_forward(ape_net* net, const float* new_points, const float* emb, const int64_t* obj,
                        int B, int N, float* r2, float* t2, void* stream);

/* Whole option-6 geometry block for B objects in one call (no host synchronisation inside):
 * PoseNet -> arg-max / pose / new cloud -> iterations x (PoseRefineNet -> fp64 compose -> next cloud).
 * Replaces pipeline/utils.py:564-571.  canonical != 0 follows DenseFusion/tools/eval_linemod.py:81-114
 * (cloud re-expressed in the composed pose every iteration); canonical == 0 reproduces pipeline/utils.py
 * as written (the refiner input is never updated, so its calls are identical: one call, one composition).
 *   poses [B,7] fp64 out (quaternion wxyz, translation); which_max [B] int32 out or NULL.             */
int ape_pose_pipeline(ape_net* estimator, ape_net* refiner, const float* out_img, int hw, const float* cloud,
                      const int64_t* choose, const int64_t* obj, int B, int N, int iterations, int canonical,
                      double* poses, int32_t* which_max, void* stream);

/* Layout of the colour-encoder output handed to the PoseNet entry points (`_ex` variants below). */
enum { APE_EMB_NCHW = 0,      /* out_img [B,32,hw]: the reference encoder's own layout (network.py:98-102)            */
       APE_EMB_NHWC = 1,      /* out_img [B,hw,32]: the same tensor in torch's channels_last memory format (one sampled
                                 point = one contiguous 128-byte line)                                                   */
       APE_EMB_GATHERED = 2 };/* out_img is emb [B,32,N], already gathered at `choose` (ape_gather_emb /
                                 ape_host_gather_*); hw and choose are ignored                                           */
int ape_posenet_forward_ex(ape_net* net, const float* out_img, int hw, int emb_layout, const float* cloud,
                           const int64_t* choose, const int64_t* obj, int B, int N, float* pred_r, float* pred_t,
                           float* pred_c, float* emb /* may be NULL with APE_EMB_GATHERED */, void* stream);
int ape_pose_pipeline_ex(ape_net* estimator, ape_net* refiner, const float* out_img, int hw, int emb_layout,
                         const float* cloud, const int64_t* choose, const int64_t* obj, int B, int N, int iterations,
                         int canonical, double* poses, int32_t* which_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Embedding hand-over for HOST-resident encoder maps (network.py:100-102 `torch.gather(emb, 2, choose)` done where the
 * map lives, so that only the sampled columns cross PCIe: 4.1 MB instead of 157 MB per batch of 64 x 500 points).
 * ape_gather_emb: gather kernel; `out_img` is a device pointer OR a pointer into mapped pinned host memory (zero-copy
 *   reads over PCIe); layout APE_EMB_NCHW or APE_EMB_NHWC; choose [B,N] int64 (device); emb [B,32,N] fp32 out (device).
 *   Indices are clamped to [0, hw).
 * ape_host_gather_begin / _wait: the same gather on the HOST for objects [obj_begin, obj_end) into a (pinned) staging
 *   buffer emb_host [B,32,N], on a persistent thread pool (`threads` <= 0: all hardware threads); begin returns at once,
 *   wait blocks until the staging buffer is complete.  One gather in flight per process.  Pure data movement: the
 *   caller then copies emb_host[obj_begin:obj_end] to the device and calls the APE_EMB_GATHERED entry points.         */
int ape_gather_emb(const float* out_img, int hw, int layout, const int64_t* choose, int B, int N, float* emb, void* stream);
int ape_host_gather_begin(const float* out_img_host, int hw, int layout, const int64_t* choose_host, int obj_begin,
                          int obj_end, int n_points, float* emb_host, int threads);
int ape_host_gather_wait(void);

/* Candidate-pose distances of the estimator loss `Loss` (DenseFusion/lib/loss.py:30-50):
 * as ape_add_metric for B = the N per-point candidate poses of an object (pass model / target with stride 0
 * to share them), plus std_out[i] = unbiased std of the per-point distances (torch.std, loss.py:50).
 * The reference materialises pred [N,M,3] and runs an N*M-query kNN for symmetric objects; this does neither. */
int ape_add_metric_std(const float* quat, const float* trans, const float* model_points, int64_t model_stride,
                       int n_model, const float* target, int64_t target_stride, int n_target,
                       const uint8_t* symmetric, int B, float* dis, float* std_out, void* stream);

/* a12. Estimator loss `Loss` (DenseFusion/lib/loss.py:12-73), forward AND backward, for the n_cand per-point candidate
 * poses of one object (the reference's batch size is 1, train.py:216): pred_r [n_cand,4] raw quaternions, pred_t
 * [n_cand,3], pred_c [n_cand], points [n_cand,3], model_points / target [n_mesh,3] fp32; symmetric != 0 selects the
 * nearest-neighbour target (the caller passes `idx in sym_list and not refine`, loss.py:40-41).
 *   loss_dis [2] out: loss = mean_i[(dis_i + 2 std_i) c_i - w log c_i] (:53) and dis of the most confident candidate (:73)
 *   which_max [1] out; dis / std_out / term [n_cand] out (per-candidate mean distance, unbiased std, loss term)
 *   d_r [n_cand,4], d_t [n_cand,3], d_c [n_cand] out: d loss / d (pred_r, pred_t, pred_c), or NULL (forward only)
 *   new_points [n_cand,3], new_target [n_mesh,3] out or NULL (:55-69); pred_out [n_cand,n_mesh,3] out or NULL (:38)
 * Neither the [N,M,3] gather nor the N*M-query kNN of the reference is materialised (unless pred_out is requested).  */
int ape_estimator_loss(const float* pred_r, const float* pred_t, const float* pred_c, const float* points,
                       const float* model_points, const float* target, int n_cand, int n_mesh, int symmetric, float w,
                       float* loss_dis, int32_t* which_max, float* dis, float* std_out, float* term, float* d_r, float* d_t,
                       float* d_c, float* new_points, float* new_target, float* pred_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Point-cloud outlier filters, open3d 0.9.0 semantics, ragged batches
 * (points [P,3] fp64 + int32 offsets [C+1], millimetres as the reference).  Replace
 * pcd.remove_radius_outlier / compute_mahalanobis_distance / remove_statistical_outlier at
 * pc_reconstruction/open3d_utils.py:158-166 and :198-211.  max_cloud_points = largest cloud of the batch.  */
/* keep[i] = 1 iff more than nb_points points (itself included) lie strictly within `radius`.          */
int ape_radius_outlier(const double* points, const int32_t* offset, int n_clouds, int max_cloud_points,
                       int nb_points, double radius, uint8_t* keep /* [P] */,
                       int32_t* n_neighbors /* [P] or NULL */, void* stream);
/* dist[i] = Mahalanobis distance to the cloud's mean (population covariance); std_out[c] = np.std(|dist|). */
int ape_mahalanobis(const double* points, const int32_t* offset, int n_clouds, double* dist /* [P] or NULL */,
                    double* std_out /* [C] or NULL */, void* stream);
/* avg_dist[i] = mean distance to the nb_neighbors nearest points (itself included); keep[i] = 1 iff
 * 0 < avg_dist[i] < mean + ratio * std (Bessel-corrected) over the cloud.  ratio = std_ratio_dev[c] when
 * the device array is given (e.g. ape_mahalanobis' std_out, as the reference feeds it), else std_ratio.   */
int ape_statistical_outlier(const double* points, const int32_t* offset, int n_clouds, int max_cloud_points,
                            int nb_neighbors, const double* std_ratio_dev, double std_ratio, uint8_t* keep,
                            double* avg_dist /* [P] out */, double* threshold /* [C] or NULL */, void* stream);
/* Ordered compaction (pcd.select_down_sample): cloud c of the result occupies
 * out_points[offset[c] : offset[c] + out_counts[c]]; out_index = kept point's index within its cloud.     */
int ape_compact_points(const double* points, const int32_t* offset, const uint8_t* keep, int n_clouds,
                       double* out_points, int32_t* out_counts, int32_t* out_index /* or NULL */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Refiner training step (a16): DenseFusion/tools/train.py:215-233 -- per sample `refiner(new_points, emb,
 * idx)` -> `criterion_refine(...)` (lib/loss_refiner.py:12-64) -> `dis.backward()`, Adam step per batch
 * (train.py:149, :231-233).  bf16 tensor-core forward / backward, fp32 master weights and gradients.
 *
 * All trainable parameters of PoseRefineNet live in ONE caller-owned flat fp32 device vector (and a flat
 * gradient vector of the same size, which is what the caller all-reduces over NCCL between ranks).
 * ape_refiner_trainer_layout returns the vector length and the offset of each of the 24 reference
 * state_dict tensors (weight, bias of feat.conv1, feat.e_conv1, feat.conv2, feat.e_conv2, feat.conv5,
 * feat.conv6, conv1_r, conv1_t, conv2_r, conv2_t, conv3_r, conv3_t) in it; shapes are the reference's.  */
typedef struct ape_trainer ape_trainer;   /* opaque */
int64_t ape_refiner_trainer_layout(int num_obj, int64_t* offsets24 /* out, may be NULL */);
int ape_refiner_trainer_create(float* params, float* grads, int num_obj, int max_batch, int max_points,
                               ape_trainer** out);
int ape_refiner_trainer_destroy(ape_trainer* tr);
/* Re-derive the bf16 weight copies from `params`; call after every change of the flat vector
 * (optimizer step, checkpoint load).                                                               */
int ape_refiner_trainer_sync_weights(ape_trainer* tr, void* stream);
/* Forward as ape_refiner_forward (network.py:187-206) in plain bf16, keeping the activations.       */
int ape_refiner_trainer_forward(ape_trainer* tr, const float* new_points, const float* emb,
                                const int64_t* obj, int B, int N, float* r2, float* t2, void* stream);
/* Backward of the forward just run (same inputs): grads += d(sum_b <d_r[b], r2[b]> + <d_t[b], t2[b]>)/dW.
 * Gradients ACCUMULATE, as repeated dis.backward() calls do in train.py:222 (zero them per step).   */
int ape_refiner_trainer_backward(ape_trainer* tr, const float* new_points, const float* emb,
                                 const int64_t* obj, int B, int N, const float* d_r, const float* d_t,
                                 void* stream);
/* The accumulation phase of one optimizer step in ONE call (train.py:215-223 for a batch): optional zeroing of the flat
 * gradient, then `iterations` x (forward -> Loss_refine fwd + bwd -> backward), cloud and target re-expressed in the
 * predicted frame in between.  points [B,N,3], emb [B,32,N], obj [B], target / model_points [B,n_mesh,3], symmetric [B]
 * u8 or NULL; dis [iterations,B] out.  No host synchronisation; graph-capturable after the first call.
 * ape_refiner_trainer_adam = ape_adam_step on the trainer's vectors + ape_refiner_trainer_sync_weights.          */
int ape_refiner_trainer_step(ape_trainer* tr, const float* points, const float* emb, const int64_t* obj,
                             const float* target, const float* model_points, const uint8_t* symmetric, int B, int N,
                             int n_mesh, int iterations, int zero_grad, float* dis, void* stream);
int ape_refiner_trainer_adam(ape_trainer* tr, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                             float eps, int step, float grad_scale, void* stream);
/* Data-parallel overlap.  The flat vectors are laid out in two contiguous blocks: [0, bulk_begin) = conv1 .. conv5 (weights
 * and biases) and [bulk_begin, total) = conv6 + the heads (89 % of the parameters).  The tail block of the gradient is
 * final as soon as the LAST iteration's conv6 weight gradient has been written; ape_refiner_trainer_step records an event
 * at that point and this call makes `side_stream` wait for it, so the caller's all-reduce of grads[bulk_begin:] issued on
 * that stream overlaps the rest of the backward pass (all-reduce grads[:bulk_begin] on the step's stream afterwards and
 * join the streams before ape_refiner_trainer_adam).  Not for use with a graph-captured step.                        */
int ape_refiner_trainer_wait_bulk(ape_trainer* tr, void* side_stream, int64_t* bulk_begin /* out, may be NULL */);
/* Loss_refine forward + backward for B objects (lib/loss_refiner.py:12-64):
 *   quat [B,4], trans [B,3], model_points [B,M,3], target [B,M,3], points [B,N,3], symmetric [B] u8 or NULL
 *   dis [B] out; d_r [B,4], d_t [B,3] = d dis[b] / d (quat, trans) out (NULL to skip);
 *   new_points [B,N,3], new_target [B,M,3] out (NULL to skip): next iteration's cloud / target (:51-60). */
int ape_refine_loss(const float* quat, const float* trans, const float* model_points, const float* target,
                    int n_mesh, const float* points, int n_points, const uint8_t* symmetric, int B,
                    float* dis, float* d_r, float* d_t, float* new_points, float* new_target, void* stream);
/* torch.optim.Adam update (train.py:149) on flat vectors; grads are multiplied by grad_scale first
 * (1/world_size after a sum all-reduce).  `step` counts from 1.                                     */
int ape_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APE_B200_H */
