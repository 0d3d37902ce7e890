/*
 * seevcn_b200 — C-ABI of the B200-native SEE-VCN object-completion + voxelization hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*; pass
 * NULL for the legacy default stream, which is what the reference launches on).  No torch
 * types cross this boundary.  The caller owns and allocates every input, output and
 * workspace buffer (reference convention: SURVEY.md §8(b) "Ownership"); nothing in here
 * allocates on the hot path.  Every function returns 0 on success and a non-zero
 * SEEVCN_E_* code on failure (the reference's launchers fprintf+exit(-1) instead,
 * e.g. detector3d/pcdet/ops/pointnet2/pointnet2_batch/src/sampling_gpu.cu:46-50);
 * seevcn_last_error() gives the message for the calling thread.
 *
 * "ref:" comments cite the reference interface each symbol replaces, relative to the
 * darrenjkt/SEE-VCN checkout.
 */
#ifndef SEEVCN_B200_H_
#define SEEVCN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEEVCN_OK            0
#define SEEVCN_E_INVALID     1   /* bad argument (shape, null pointer, unsupported size)   */
#define SEEVCN_E_CUDA        2   /* CUDA runtime / launch error                            */
#define SEEVCN_E_WORKSPACE   3   /* workspace too small                                    */
#define SEEVCN_E_UNSUPPORTED 4   /* device is not sm_100                                   */

typedef void* seevcn_stream_t;   /* cudaStream_t */

int         seevcn_abi_version(void);
const char* seevcn_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long seevcn_launch_count(void);
/* Optional timing of the library's launch groups with CUDA events recorded on the stream each group is launched
 * on (the reference has no profiling hooks).  seevcn_prof_enable(1) clears and starts recording,
 * (0) stops; returns the previous state.  seevcn_prof_report synchronises the recorded events and writes one
 * line per group name, "<name> <launch groups> <total ms>\n", into buf (NUL-terminated), then clears. */
int         seevcn_prof_enable(int on);
int         seevcn_prof_report(char* buf, size_t cap);
/* Returns 0 when device `dev` is an sm_100 part this library was built for. */
int         seevcn_check_device(int dev);

/* ---------------------------------------------------------------- stage 1: crop ---- */

/* ref: void points_in_boxes_launcher(int batch_size, int boxes_num, int pts_num,
 *          const float *boxes, const float *pts, int *box_idx_of_points)
 *      detector3d/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:339-359
 *      (kernel :313-336, predicate :16-36), bound by roiaware_pool3d.cpp:98-118.
 * boxes (B,T,7) [x,y,z,dx,dy,dz,heading], pts (B,P,3), out (B,P) int32: lowest box index
 * containing the point, else -1.  Unlike the reference the output does NOT need to be
 * pre-filled with -1: every element is written. */
int seevcn_points_in_boxes(int batch_size, int boxes_num, int pts_num,
                           const float* boxes, const float* pts,
                           int* box_idx_of_points, seevcn_stream_t stream);

/* ref: int points_in_boxes_cpu(at::Tensor boxes, at::Tensor pts, at::Tensor pts_indices)
 *      detector3d/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168
 *      (MARGIN = 1e-2, dense (T,P) 0/1 matrix, no first-box-wins).
 * boxes (T,7), pts (P,3), out (T,P) int32.  Runs on the GPU; "cpu" in the reference name
 * describes where the reference computes it, not the semantics. */
int seevcn_points_in_boxes_dense(int boxes_num, int pts_num,
                                 const float* boxes, const float* pts,
                                 int* pts_indices, seevcn_stream_t stream);
/* Same, with box_trig (T,2) = host-computed {cosf(-heading), sinf(-heading)} per box so
 * the result is bit-identical to the reference's x86 build (glibc cosf/sinf differ from
 * CUDA's in the last ulp).  box_trig == NULL computes them on the device. */
int seevcn_points_in_boxes_dense_trig(int boxes_num, int pts_num,
                                      const float* boxes, const float* box_trig, const float* pts,
                                      int* pts_indices, seevcn_stream_t stream);

/* Crop with compaction (the list form SEE's per-object crop needs:
 * see/surface_completion/SEE_VCN.py:61-82 does one crop per box on the host).
 * In addition to box_idx_of_points (B,P) it emits, per (frame, box), the ascending list of
 * point indices inside the box:
 *   box_counts  (B,T)   int32  number of points in each box
 *   box_offsets (B,T)   int32  exclusive prefix of box_counts within the frame
 *   box_points  (B,P)   int32  point indices grouped by box (first sum(counts) entries used)
 * workspace: seevcn_crop_workspace_bytes(B,T,P) bytes. */
size_t seevcn_crop_workspace_bytes(int batch_size, int boxes_num, int pts_num);
int seevcn_crop_points_in_boxes(int batch_size, int boxes_num, int pts_num,
                                const float* boxes, const float* pts,
                                int* box_idx_of_points, int* box_counts, int* box_offsets,
                                int* box_points, void* workspace, size_t workspace_bytes,
                                seevcn_stream_t stream);

/* ref: ResamplePoints.__call__  see/surface_completion/models/vcn/datasets/data_transforms.py:247-262
 * (tile to >= n_points, take the first n_points of a random permutation).  The permutation
 * is supplied by the caller as `choice` (num_obj, n_points) int32 indices into the tiled
 * list, so a seeded host RNG reproduces the reference bit for bit.
 * For object o: src = box_points[frame_of[o]*P + box_offsets[o] + choice[o,j] % count[o]],
 * out[o,j,:] = pts[frame_of[o], src, :].
 * obj_frame/obj_box (num_obj) name the (frame, box) each object came from. */
int seevcn_resample_gather(int num_obj, int n_points, int boxes_num, int pts_num,
                           const float* pts, const int* box_counts, const int* box_offsets,
                           const int* box_points, const int* obj_frame, const int* obj_box,
                           const int* choice, float* out, seevcn_stream_t stream);

/* ref: SEE_VCN.isolate_gt_pts keeps the boxes with at least MIN_LIDAR_PTS points (see/surface_completion/SEE_VCN.py:71).
 * box_counts (batch, boxes_num) int32 -> the (frame, box) pairs with count >= min_pts in frame-major order (the order of
 * numpy.argwhere on the host copy of the counts): obj_frame / obj_box (capacity batch*boxes_num) int32, num_obj (1) int32.
 * Lets the host, which only needs the NUMBER of objects to size the launches, skip uploading the list. */
int seevcn_select_objects(int batch, int boxes_num, const int* box_counts, int min_pts,
                          int* obj_frame, int* obj_box, int* num_obj, seevcn_stream_t stream);

/* Same draw made on the device: choice[o,j] = perm_o(j), the first n_points entries of a pseudo-random
 * permutation of the tiled list (4-round Feistel network + cycle walking, keyed by seed and the object's
 * frame*T+box), so no host RNG, no `choice` upload.  seevcn_resample_perm() evaluates the same permutation
 * on the host (tests). */
int seevcn_resample_gather_rng(int num_obj, int n_points, int boxes_num, int pts_num, unsigned seed,
                               const float* pts, const int* box_counts, const int* box_offsets,
                               const int* box_points, const int* obj_frame, const int* obj_box,
                               float* out, seevcn_stream_t stream);
unsigned seevcn_resample_perm(unsigned j, unsigned n, unsigned seed, unsigned frame_box);

/* ---------------------------------------------------- stage 3: furthest point sampling */

/* ref: void farthest_point_sampling_kernel_launcher(int b, int n, int m,
 *          const float *dataset, float *temp, int *idxs)
 *      detector3d/pcdet/ops/pointnet2/pointnet2_batch/src/sampling_gpu.cu:218-260 (kernel :100-216)
 * dataset (B,N,3) f32, idxs (B,M) int32.  temp (B,N) may be NULL; when given it receives
 * the final min-distance array exactly as the reference leaves it (it does not need the
 * 1e10 pre-fill the reference's python wrapper does). */
int seevcn_furthest_point_sampling(int b, int n, int m, const float* dataset,
                                   float* temp, int* idxs, seevcn_stream_t stream);

/* ---------------------------------------------------- stage 4: gather / group / kNN -- */

/* ref: gather_points_kernel_launcher_fast  sampling_gpu.cu:33-51.
 * points (B,C,N), idx (B,M) -> out (B,C,M) */
int seevcn_gather_points(int b, int c, int n, int npoints, const float* points,
                         const int* idx, float* out, seevcn_stream_t stream);

/* ref: group_points_kernel_launcher_fast  pointnet2_batch/src/group_points_gpu.cu:75-92.
 * points (B,C,N), idx (B,P,S) -> out (B,C,P,S) */
int seevcn_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                        const int* idx, float* out, seevcn_stream_t stream);

/* k nearest neighbours, ascending.  Semantics of scipy cKDTree.query(k) / dist.topk(k,
 * largest=False) as used by see/surface_completion/models/vcn/utils/sampling.py:30-34,59-61.
 * ref_pts (B,R,3), query (B,Q,3) -> dist (B,Q,k) f32 Euclidean (not squared), idx (B,Q,k) int32.
 * 1 <= k <= 64, k <= R.  dist may be NULL. */
int seevcn_knn(int b, int r, int q, int k, const float* ref_pts, const float* query,
               float* dist, int* idx, seevcn_stream_t stream);

/* ref: partial_with_KDTree / get_partial_mesh_batch  sampling.py:8-41,69-80.
 * For every object: S = union of the k nearest `complete` points of every `partial`
 * point; out = complete[sorted(S)] repeated cyclically to surface_pts rows.
 * partial (B,Np,3), complete (B,R,3) -> out (B,surface_pts,3), sel_count (B) int32 = |S|.
 * R <= 16384, Np <= 4096.  workspace: seevcn_knn_surface_select_workspace_bytes(b, n_partial, r) bytes of
 * 16-byte aligned device scratch (Morton-ordered copies of the clouds, per-block bounding boxes, the union bit masks). */
size_t seevcn_knn_surface_select_workspace_bytes(int b, int n_partial, int r);
int seevcn_knn_surface_select(int b, int n_partial, int r, int k, int surface_pts,
                              const float* partial, const float* complete,
                              float* out, int* sel_count,
                              void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* ref: get_largest_cluster(_batch)  see/surface_completion/models/vcn/utils/sampling.py:83-109
 * (open3d cluster_dbscan(eps, min_points) -> largest cluster -> tiled to total_pts rows), called with
 * min_points = 2 at see/surface_completion/models/VCN.py:95-98.  min_points in {1, 2}: DBSCAN is then exactly
 * connected components of the eps-graph (strict d < eps, float64), isolated points are noise when
 * min_points = 2.  pts (B,n,3), n <= 8192 -> out (B,total_pts,3): members of the largest component in
 * ascending row order, repeated cyclically; out_count (B) = member rows (0: everything was noise, rows zero);
 * out_distinct (B) or NULL = how many of the clustered rows are members, i.e. the leading rows of `out` before the
 * cyclic repetition starts (== out_count unless the input was tiled and clustered through the periodic entry).
 * PARITY UNPINNED (open3d is not vendored). */
int seevcn_largest_cluster(int b, int n, int total_pts, double eps, int min_points, const float* pts,
                           float* out, int* out_count, int* out_distinct, seevcn_stream_t stream);

/* Same result, for clouds the caller knows to be tiled: period (B) int32 DEVICE, pts[b][r] == pts[b][r % period[b]]
 * (the output of seevcn_knn_surface_select with period = sel_count).  Only the period[b] distinct rows are
 * clustered, each weighted by its multiplicity (out_count is the weighted size, as np.bincount sees it on the tiled
 * cloud); out_distinct is the number of distinct member rows — what np.unique(clustered) keeps (SEE_VCN.py:113,244) and
 * what the splice / voxelization stages take as the object's row count.  eps > 0. */
int seevcn_largest_cluster_periodic(int b, int n, int total_pts, double eps, int min_points, const float* pts,
                                    const int* period, float* out, int* out_count, int* out_distinct,
                                    seevcn_stream_t stream);

/* ------------------------------------------- stages 2+5: VCN forward (canonicalise+MLP) */

/* Folded fp32 parameters, all DEVICE pointers, row-major (out,in) like torch Linear /
 * Conv1d(k=1) weights.  BatchNorm (eval) must already be folded into the preceding
 * conv's weight and bias by the caller (python host does it).  Names follow the
 * reference state-dict (see/surface_completion/models/vcn/models/VCN_VC.py:116-131):
 *   pose_encoder.{0,2,4}  3->64->128->1024   (LeakyReLU 0.01 after 0 and 2)   [VC only]
 *   pose_fc.{0,2}         1024->512->9       (LeakyReLU after 0)             [VC only]
 *   encoder.mlp_conv1.{0(+bn1),3}  3->128->256
 *   encoder.mlp_conv2.{0(+bn1),3}  512->512->1024
 *   shape_fc.{0,2,4}      1024->1024->1024->3*num_coarse
 * For VCN_CN (VCN_CN.py:111-157) the pose_* pointers are NULL. */
typedef struct seevcn_vcn_params {
    const float *pose_enc0_w, *pose_enc0_b;   /* (64,3)      (64)   */
    const float *pose_enc2_w, *pose_enc2_b;   /* (128,64)    (128)  */
    const float *pose_enc4_w, *pose_enc4_b;   /* (1024,128)  (1024) */
    const float *pose_fc0_w,  *pose_fc0_b;    /* (512,1024)  (512)  */
    const float *pose_fc2_w,  *pose_fc2_b;    /* (9,512)     (9)    */
    const float *enc1_0_w,    *enc1_0_b;      /* (128,3)     (128)  BN folded */
    const float *enc1_3_w,    *enc1_3_b;      /* (256,128)   (256)  */
    const float *enc2_0_w,    *enc2_0_b;      /* (512,512)   (512)  BN folded */
    const float *enc2_3_w,    *enc2_3_b;      /* (1024,512)  (1024) */
    const float *fc0_w,       *fc0_b;         /* (1024,1024) (1024) */
    const float *fc2_w,       *fc2_b;         /* (1024,1024) (1024) */
    const float *fc4_w,       *fc4_b;         /* (3*num_coarse,1024) (3*num_coarse) */
    int num_coarse;                           /* 1024 in the shipped models */
    int viewer_centred;                       /* 1 = VCN_VC, 0 = VCN_CN */
} seevcn_vcn_params;

typedef struct seevcn_vcn_model seevcn_vcn_model;   /* opaque: packed bf16 weights on device */

/* Packs the parameters into the kernels' bf16 operand layouts (allocates device memory —
 * model-load time, not the hot path). */
int  seevcn_vcn_create(const seevcn_vcn_params* params, seevcn_vcn_model** out_model,
                       seevcn_stream_t stream);
void seevcn_vcn_destroy(seevcn_vcn_model* model);

size_t seevcn_vcn_workspace_bytes(const seevcn_vcn_model* model, int num_obj, int n_pts);

/* ref: VCN_VC.forward  see/surface_completion/models/vcn/models/VCN_VC.py:178-213
 *      VCN_CN.forward  see/surface_completion/models/vcn/models/VCN_CN.py:142-157
 * input (B,N,3) f32; gt_boxes (B,7) f32 (VCN_CN only, else NULL)
 * -> coarse (B,num_coarse,3) f32, reg_rot (B,3,3) f32, reg_centre (B,3) f32 (VC only; may be NULL).
 * precision: 0 = bf16 operands / fp32 accumulate on tcgen05 with the per-point layers fused into chains whose
 * activations stay in tensor memory (vcn_chain.cu); 1 = fp32 SIMT (validation); 2 = the bf16 tcgen05 path with one GEMM
 * launch per layer (vcn_tc.cu; same layers and roundings as 0 up to accumulation order, kept for cross-checking). */
int seevcn_vcn_forward(const seevcn_vcn_model* model, int num_obj, int n_pts,
                       const float* input, const float* gt_boxes,
                       float* coarse, float* reg_rot, float* reg_centre,
                       void* workspace, size_t workspace_bytes, int precision,
                       seevcn_stream_t stream);

/* One shared-MLP layer on the tcgen05 path, standalone (the building block of seevcn_vcn_forward;
 * ref: nn.Conv1d(k=1) / nn.Linear as used in VCN_VC.py:116-131):
 *   Y[r, c] = act(sum_k bf16(X[r,k]) * bf16(W[c,k]) + bias[c] + obj_bias[r / rows_per_obj, c])   (fp32 accumulate)
 * X (rows,cin) f32, W (cout,cin) f32, bias (cout) or NULL, obj_bias (rows/rows_per_obj, cout) or NULL,
 * act 0 none / 1 ReLU / 2 LeakyReLU(0.01).  Y (rows,cout) f32 or NULL; colmax (rows/rows_per_obj, cout) f32 or
 * NULL receives the per-object max over rows (must be pre-filled with -inf).  */
size_t seevcn_linear_bf16_workspace_bytes(int rows, int cin, int cout);
int seevcn_linear_bf16(int rows, int cin, int cout, const float* X, const float* W, const float* bias,
                       const float* obj_bias, int rows_per_obj, int act, float* Y, float* colmax,
                       void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* ------------------------------------------------ mask-based isolation */

/* ref: CustomDatasetObjects.map_pointcloud_to_image
 *      see/surface_completion/datasets/custom_dataset/custom_dataset_objects.py:141-193 (float64 numpy on the host).
 * pts (N,3) f32 DEVICE; lidar2cam[12] = rows of extrinsic[:3,:], intrinsic[9] = 3x3 row major, distcoeff[5]: HOST float64
 * (calibration constants).  equidistant != 0 selects the fisheye model, else pinhole (k1 k2 p1 p2 k3 = distcoeff[0..4]).
 * -> uv (N,2) int32 = (np.round(u), np.round(v)) for points in the field of view, (-1,-1) otherwise; fov (N) uint8;
 * depth (N) f32 or NULL (camera-frame z). */
int seevcn_project_points(int num_points, const float* pts, const double* lidar2cam, const double* intrinsic,
                          const double* distcoeff, int equidistant, int img_w, int img_h,
                          int* uv, unsigned char* fov, float* depth, seevcn_stream_t stream);

/* ref: get_pts_in_mask  see/surface_completion/datasets/shared_utils.py:36-106 (mask[v, u] lookup per instance).
 * masks (I, img_h, img_w) uint8 binary instance masks (the polygon -> mask rasterisation, pycocotools annToMask, stays
 * with the caller) -> lists (I, N) int32: ascending indices of the in-view points inside mask i; counts (I). */
int seevcn_points_in_masks(int num_points, int num_inst, int img_w, int img_h, const int* uv, const unsigned char* fov,
                           const unsigned char* masks, int* lists, int* counts, seevcn_stream_t stream);

/* ref: SEE_VCN.isolate_det_pts  see/surface_completion/SEE_VCN.py:144-181: per instance with more than
 * min_instance_pts points, eps = clip(eps_scaling * |centre| * tan(vres deg), min_eps, max_eps) (adaptive != 0) or the
 * fixed eps, open3d cluster_dbscan(eps, min_points), the largest cluster (first on ties), kept when it holds more than
 * min_instance_pts points.  lists (I, stride) indices into pts (P,3) with counts (I) (the output of
 * seevcn_points_in_masks) -> out_lists (I, stride): the cluster's point indices, ascending; out_counts (I): their number
 * (0: dropped); out_eps (I) float64 or NULL.  Instances of up to 9600 points are clustered in shared memory; larger ones
 * take 24 B per point from `workspace` (may be NULL) and report -1 when it is too small — call again with more.
 * PARITY UNPINNED (open3d). */
int seevcn_dbscan_largest(int num_inst, int stride, const float* pts, const int* lists, const int* counts,
                          int adaptive, double eps, double vres_deg, double eps_scaling, double min_eps, double max_eps,
                          int min_points, int min_instance_pts, int* out_lists, int* out_counts, double* out_eps,
                          void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* ref: ResamplePoints (data_transforms.py:247-262) applied to isolated instances (models/VCN.py:52-53): object o is
 * list row obj_inst[o]; out[o, j] = pts[lists[inst][perm(j) % count]], perm = seevcn_resample_perm(j, reps * count, seed,
 * inst).  out (num_obj, n_points, 3). */
int seevcn_resample_lists(int num_obj, int n_points, int stride, unsigned seed, const float* pts, const int* lists,
                          const int* counts, const int* obj_inst, float* out, seevcn_stream_t stream);

/* ---------------------------------------------- splice: completed clouds replace raw points */

/* ref: SEE_VCN.replace_with_completed_pts  see/surface_completion/SEE_VCN.py:247-265 (demo twin
 * demo/see_vcn_dataset.py:127-135): a raw frame point is dropped when its nearest completed point is closer than
 * thresh (`compute_point_cloud_distance(...) < point_dist_thresh`, float64), the rest is stacked under the completed
 * points.  PARITY UNPINNED (open3d is not vendored).
 *   frame_pts (F,P,3) f32; obj_pts (O,S,3) f32 completed clouds, object o belongs to frame obj_frame[o] (O int32,
 *   NON-DECREASING) and only its first obj_count[o] rows count (O int32, NULL = S; 0 = object contributes nothing).
 *   keep (F,P) uint8: 1 = the point survives.
 *   merged (F,out_stride,3) f32 or NULL: per frame [rows of its objects in object order ++ surviving points in point
 *   order]; merged_count (F) int32 rows written per frame (out_stride >= P + the frame's object rows);
 *   completed_count (F) int32 or NULL = how many of them are object rows.
 *   frame_rows (F, frame_rows_stride, 3) + frame_row_count (F) int32, or NULL: the completed rows of every frame as one
 *   pre-merged block that replaces the per-object rows at the head of the merged cloud — pass the output of
 *   seevcn_unique_rows_frames to get the reference's `vstack(np.unique(all object rows), pcd_without_object)`
 *   (SEE_VCN.py:244,262: lexicographic order, cross-object duplicates removed) bit for bit.  The keep mask does not
 *   depend on it.
 * workspace: seevcn_splice_workspace_bytes(F, P, O) bytes, 16-byte aligned. */
size_t seevcn_splice_workspace_bytes(int num_frames, int pts_per_frame, int num_obj);
int seevcn_splice(int num_frames, int pts_per_frame, const float* frame_pts,
                  int num_obj, int pts_per_obj, const float* obj_pts, const int* obj_count, const int* obj_frame,
                  double thresh, unsigned char* keep,
                  int out_stride, float* merged, int* merged_count, int* completed_count,
                  const float* frame_rows, int frame_rows_stride, const int* frame_row_count,
                  void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* ref: sc_model_ret['all_instances'] = np.unique(np.vstack(sc_model_ret['clustered']), axis=0)
 *      see/surface_completion/SEE_VCN.py:113,244.  Per frame: the first obj_count[o] rows (NULL = all) of the frame's objects
 * (obj_frame non-decreasing), sorted lexicographically by (x, y, z) with duplicates removed.
 * -> uniq (F, out_stride, 3) f32, ucount (F) int32; out_stride >= the rows any frame's objects hold together;
 * pts_per_obj <= 4096.  workspace: seevcn_unique_rows_frames_workspace_bytes(F, O, S, out_stride). */
size_t seevcn_unique_rows_frames_workspace_bytes(int num_frames, int num_obj, int pts_per_obj, int out_stride);
int seevcn_unique_rows_frames(int num_frames, int num_obj, int pts_per_obj, const float* obj_pts, const int* obj_count,
                              const int* obj_frame, int out_stride, float* uniq, int* ucount,
                              void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* ------------------------------------------------------------ stage 6: voxelization -- */

/* ref: MeanVFE.forward  detector3d/pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31
 * voxels (M,T,C) f32, voxel_num_points (M) f32 -> voxel_features (M,C) f32
 * = sum over T / max(num_points, 1). */
int seevcn_mean_vfe(int num_voxels, int max_points, int num_features, const float* voxels,
                    const float* voxel_num_points, float* voxel_features,
                    seevcn_stream_t stream);
/* Same with the int32 counts the voxel generator produces (the reference casts them to float on the way to the GPU,
 * detector3d/pcdet/models/__init__.py:34). */
int seevcn_mean_vfe_int(int num_voxels, int max_points, int num_features, const float* voxels,
                        const int* voxel_num_points, float* voxel_features,
                        seevcn_stream_t stream);

/* ref: DynamicMeanVFE.forward  detector3d/pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:37-76
 * points (N,1+C) f32 rows [batch_idx,x,y,z,...]; pc_range[6], voxel_size[3], grid_size[3]
 * are HOST arrays (grid constants, like the python attributes of the reference module); batch_size is
 * batch_dict['batch_size'] (>= 1; rows whose batch index lies outside [0, batch_size) are ignored).
 * Outputs (capacity max_voxels rows each):
 *   voxel_coords   (M,4) int32 [b,z,y,x]
 *   voxel_features (M,C) f32   mean of all in-voxel points
 *   voxel_counts   (M)   int32 (the reference computes unq_cnt and drops it, :63)
 *   num_voxels     (1)   int32 (device)  M; rows beyond max_voxels are dropped, M still reports all of them
 * Rows are ALWAYS ordered by the reference's merge key b*XYZ + x*YZ + y*Z + z, i.e. torch.unique order (64-bit, so
 * batch >= 24 on the Waymo grid does not overflow as the reference's int32 key does); `sorted` is kept for source
 * compatibility and ignored.  No sort and no hash table: points are bucketed by the high key bits (counting pass +
 * look-back scan + scatter) and a warp per bucket ranks its voxels with an occupancy bitmap.  xyz means are accumulated
 * as integers relative to the voxel origin: bit-reproducible run to run, equal to the float64 mean rounded to fp32.
 * workspace: seevcn_dynamic_voxelize_workspace_bytes(N, C, batch_size, grid_size) bytes (0: batch x grid too large). */
size_t seevcn_dynamic_voxelize_workspace_bytes(int num_points, int num_features, int batch_size, const int* grid_size);
int seevcn_dynamic_voxelize(int num_points, int num_features, const float* points,
                            const float* pc_range, const float* voxel_size, const int* grid_size,
                            int max_voxels, int sorted, int batch_size,
                            int* voxel_coords, float* voxel_features, int* voxel_counts,
                            int* num_voxels, void* workspace, size_t workspace_bytes,
                            seevcn_stream_t stream);

/* The frame pipeline's form of the same op: rows come from two device arrays — the raw frames frame_pts (F,P,3),
 * batch index = frame, and the completed object clouds obj_pts (O,S,3), batch index obj_frame[o] (O int32) — so the
 * [batch_idx,x,y,z] matrix the reference concatenates on the host (detector3d/pcdet/datasets/dataset.py:187-192
 * after SEE_VCN.py:247-265 merged the completed points into the frame) is never materialised.  C = 3.
 * workspace: seevcn_dynamic_voxelize_workspace_bytes(F*P + O*S, 3, F, grid_size). */
int seevcn_dynamic_voxelize_frames(int num_frames, int pts_per_frame, const float* frame_pts,
                                   int num_obj, int pts_per_obj, const float* obj_pts, const int* obj_frame,
                                   const float* pc_range, const float* voxel_size, const int* grid_size,
                                   int max_voxels, int sorted,
                                   int* voxel_coords, float* voxel_features, int* voxel_counts, int* num_voxels,
                                   void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* Same, after the splice step: frame point (f,p) is skipped when frame_keep[f*P+p] == 0 (frame_keep (F,P) uint8 from
 * seevcn_splice, NULL = keep all) and object o contributes only its first obj_count[o] rows (obj_count (O) int32 DEVICE,
 * NULL = all pts_per_obj rows; the distinct rows of a cyclically tiled cloud, i.e. sel_count / out_count of the stages
 * above) — the rows of the reference's merged frame cloud `vstack(all_instances, pcd_without_object)`
 * (SEE_VCN.py:244,247-265). */
int seevcn_dynamic_voxelize_spliced(int num_frames, int pts_per_frame, const float* frame_pts,
                                    const unsigned char* frame_keep,
                                    int num_obj, int pts_per_obj, const float* obj_pts, const int* obj_frame,
                                    const int* obj_count,
                                    const float* pc_range, const float* voxel_size, const int* grid_size,
                                    int max_voxels, int sorted,
                                    int* voxel_coords, float* voxel_features, int* voxel_counts, int* num_voxels,
                                    void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* ref: VoxelGeneratorWrapper.generate  detector3d/pcdet/datasets/processor/data_processor.py:44-60
 * (spconv hard voxelization: first-seen voxel order, first max_points points per voxel in
 * point order, at most max_voxels voxels).  points (N,C) f32 (xyz first).
 *   voxels (max_voxels,max_points,C) f32 zero padded, coordinates (max_voxels,3) int32 zyx,
 *   num_points_per_voxel (max_voxels) int32, num_voxels (1) int32 (device).
 * PARITY UNPINNED: spconv is not vendored in the reference (docker/Dockerfile:58). */
size_t seevcn_hard_voxelize_workspace_bytes(int num_points, int max_points, int max_voxels);
int seevcn_hard_voxelize(int num_points, int num_features, const float* points,
                         const float* pc_range, const float* voxel_size, const int* grid_size,
                         int max_points, int max_voxels,
                         float* voxels, int* coordinates, int* num_points_per_voxel,
                         int* num_voxels, void* workspace, size_t workspace_bytes,
                         seevcn_stream_t stream);

/* ref: DataProcessor.mask_points_and_boxes_outside_range  detector3d/pcdet/datasets/processor/data_processor.py:78-91 with
 * common_utils.mask_points_by_range (pcdet/utils/common_utils.py:60-63): keeps the rows with limit_range[0] <= x <=
 * limit_range[3] and limit_range[1] <= y <= limit_range[4], in order.  points (N,C) -> out (N,C) (first *out_count rows),
 * out_count (1) int32 DEVICE.  limit_range[6] HOST.  workspace: seevcn_mask_points_by_range_workspace_bytes(N). */
size_t seevcn_mask_points_by_range_workspace_bytes(int num_points);
int seevcn_mask_points_by_range(int num_points, int num_features, const float* points, const float* limit_range,
                                float* out, int* out_count, void* workspace, size_t workspace_bytes, seevcn_stream_t stream);
/* ref: DataProcessor.shuffle_points  data_processor.py:93-103 (points[np.random.permutation(n)]): out[j] =
 * points[perm(j)], perm a seeded bijection of [0, n) evaluated on the fly; seevcn_shuffle_perm gives it on the host. */
int seevcn_shuffle_points(int num_points, int num_features, unsigned seed, const float* points, float* out,
                          seevcn_stream_t stream);
unsigned seevcn_shuffle_perm(unsigned j, unsigned n, unsigned seed);

/* The batched form the frame pipeline uses (the reference voxelizes frame by frame in DataLoader workers and pads the
 * per-frame coordinates with the batch index when collating, detector3d/pcdet/datasets/dataset.py:193-198):
 * points (F, stride, C) f32, row q of frame f takes part iff q < counts[f] (counts (F) int32 DEVICE, NULL = all rows).
 * Per frame: first-seen voxel order, cap max_voxels, first max_points points per voxel.  Outputs, padded per frame:
 *   voxels (F, max_voxels, max_points, C), coordinates (F, max_voxels, 4) int32 [frame, z, y, x],
 *   num_points_per_voxel (F, max_voxels) int32, num_voxels (F) int32 (device); slots >= num_voxels[f] are undefined.
 * The x/y range mask of data_processor.py:78-91 is implied: the grid spans POINT_CLOUD_RANGE, so a point outside the
 * range is outside the grid.  workspace: seevcn_hard_voxelize_frames_workspace_bytes(F, stride, max_points, max_voxels). */
size_t seevcn_hard_voxelize_frames_workspace_bytes(int num_frames, int stride, int max_points, int max_voxels);
int seevcn_hard_voxelize_frames(int num_frames, int stride, int num_features, const float* points, const int* counts,
                                const float* pc_range, const float* voxel_size, const int* grid_size,
                                int max_points, int max_voxels,
                                float* voxels, int* coordinates, int* num_points_per_voxel, int* num_voxels,
                                void* workspace, size_t workspace_bytes, seevcn_stream_t stream);

/* Copies `bytes` (multiple of 4) from device memory into PINNED host memory (cudaHostAlloc / torch pin_memory: mapped
 * into the device address space) with SM stores on `stream` — for the few words the host needs while the GPU keeps
 * running (box counts, number of voxels); unlike cudaMemcpyAsync it does not queue behind bulk downloads on the copy
 * engine.  Visible to the host once an event recorded after it on `stream` has completed. */
int seevcn_copy_to_pinned(const void* src_device, void* dst_pinned_host, size_t bytes, seevcn_stream_t stream);
/* The reverse: a few words (e.g. the object list the host derived from the box counts) from PINNED host memory into
 * device memory with SM loads; the pinned buffer must stay untouched until an event recorded after it has completed. */
int seevcn_copy_from_pinned(const void* src_pinned_host, void* dst_device, size_t bytes, seevcn_stream_t stream);

/* ------------------------------------------------- end-of-path collection (row e) ---- */
/* Replaces the pickled all_gather of per-rank results (pcdet/utils/commu_utils.py:50-111, used by sc_multiproc.py:81-85).
 * Packs one batch into a contiguous record in `send`:
 *   header_words int32 {num_objects, num_voxels, frame_offset, feat_width, 0...} | clustered (clustered_words fp32) |
 *   coords (num_voxels x 4 int32) | feats (num_voxels x 3 fp32) | nums (num_voxels int32).  One kernel on `stream`. */
int seevcn_gather_pack(int header_words, int num_objects, int num_voxels, int frame_offset, int feat_width,
                       long long clustered_words, const float* clustered, const int* coords, const float* feats,
                       const int* nums, void* send, seevcn_stream_t stream);
/* Delivers `bytes` of `send` to every dst[r] (this rank's slot in the receive buffer of rank r, mapped into this process
 * by the caller: CUDA VMM / symmetric memory) with one device-to-device copy each on `stream`. */
int seevcn_gather_broadcast(const void* send, size_t bytes, void* const* dst, int num_dst, seevcn_stream_t stream);

/* ------------------------------------------------------------------ parity metric ---- */

/* ref: chamfer_dist_kernel  see/surface_completion/models/vcn/extensions/chamfer_dist/chamfer.cu:15-145
 * xyz1 (B,N,3), xyz2 (B,M,3) -> dist1 (B,N) squared distance to nearest in xyz2, dist2 (B,M). */
int seevcn_chamfer(int b, int n, int m, const float* xyz1, const float* xyz2,
                   float* dist1, float* dist2, seevcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SEEVCN_B200_H_ */
