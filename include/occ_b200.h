/*
 * occ_b200.h -- C ABI of libocc_b200.so: the B200 (sm_100a) point -> occupancy hot path.
 *
 * Drop-in boundary for the hot path of Ghostish/ObjectCentricOccCompletion.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference repo).
 * Plain pointers and sizes only; no torch types.  Unless stated otherwise:
 *   - all data pointers are DEVICE pointers on the current CUDA device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - calls are asynchronous on `stream`; nothing synchronises unless documented;
 *   - return value 0 = success, non-zero = error (see occb200_last_error()).
 */
#ifndef OCC_B200_H_
#define OCC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library -------------------------------------------------------------------------- */

/* ABI version; bumped on any layout change of the structs below. */
int occb200_abi_version(void);
/* sizeof of occb200_pose_t, occb200_sensor_t, occb200_annotate_args_t, occb200_ri_desc_t as this library was
 * compiled: a binding checks its own struct declarations against these before the first call. */
void occb200_struct_sizes(int64_t *out4);
/* Message of the last failing call on this host thread (never NULL). */
const char *occb200_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t occb200_launch_count(void);

/* Optional per-kernel timing of occb200_annotate_batch for roofline reports: while enabled, CUDA
 * events are recorded around each pipeline kernel on the caller's stream.  occb200_profile_read
 * synchronises those events and returns, per kernel kind (0 k_tracklet_presetup, 1 k_tracklet_setup + redo,
 * 2 k_scan_chunks, 3 k_frame_voxelize, 4 k_visibility_{fast,f64}, 5 side stream: k_table_setup + k_pyr_scan +
 * k_pyr_build, 6 k_visibility_recheck, 7 k_pair_build), the summed milliseconds and
 * launch counts since the last read.
 * Both arrays have occb200_profile_kinds() entries (HOST). */
void occb200_profile_enable(int on);
int occb200_profile_kinds(void);
int occb200_profile_read(double *ms_per_kind, int64_t *launches_per_kind);
/* Self-test of the f32 arctangents of the fast visibility kernel over n pseudo-random pairs (radians):
 * max_err_host[0] = max |atan2_fast(y,x) - atan2(y,x)| (full quadrant; the margins assume 2e-6),
 * max_err_host[1] = the same for the narrow path (x > 0, |y| <= x; the margins assume 1e-6).
 * Synchronises `stream`. */
int occb200_selftest_atan2(int64_t n, uint64_t seed, double *max_err_host, void *stream);

/* reduce_t of mmdet3d/ops/voxel/src/scatter_points_cuda.cu:7 */
enum { OCCB200_SUM = 0, OCCB200_MEAN = 1, OCCB200_MAX = 2 };

/* per-tracklet status of occb200_annotate_batch (what the reference does in that case) */
enum {
  OCCB200_OK = 0,
  OCCB200_SKIP_SHORT = 1,         /* len(trk) < 10: returns silently     (tools/occ/occ_annotate.py:344)  */
  OCCB200_NO_POINTS = 2,          /* AssertionError "no points"          (occ_annotate.py:129, 356-359)   */
  OCCB200_EMPTY_AFTER_FILTER = 3, /* max() of an empty tensor raises     (occ_annotate.py:433-435)        */
  OCCB200_INDEX_ERROR = 4,        /* quantised coord < -dims: IndexError (occ_annotate.py:436)            */
  OCCB200_WORK_OVERFLOW = 5       /* internal: the ray-cast work list did not fit args.items_cap (a host bound
                                     was violated); the tracklet's labels are NOT valid                    */
};

/* ---- A1: points in boxes -------------------------------------------------------------- */

/*
 * Replaces roiaware_pool3d_ext.points_in_boxes_gpu
 *   (mmdet3d/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:125-135, kernel
 *    src/points_in_boxes_cuda.cu:51-77; python wrapper points_in_boxes.py:6-50).
 * boxes f32 [B,T,7] (x,y,z_bottom,w,l,h,rz), pts f32 [B,M,3] -> out int32 [B,M]:
 * index of the first box containing the point, else -1 (the kernel writes every element;
 * the reference wrapper's -1 pre-fill is not needed).
 * trig: optional f32 [B,T,2] = cosf/sinf(float(rz + pi/2)) computed by the caller; NULL = computed on device.
 * contract: 0 = the unfused products of points_in_boxes_cpu.cpp:27-28 (with host libm trig: bit-identical to
 * points_in_boxes_cpu); 1 = the FMA contraction nvcc applies to points_in_boxes_cuda.cu:31-32 (with device
 * trig: bit-identical to the reference's CUDA kernel).
 */
int occb200_points_in_boxes_gpu(const float *boxes, const float *pts, const float *trig, int32_t *out,
                                int B, int T, int M, int contract, void *stream);

/* Replaces roiaware_pool3d_ext.points_in_boxes_batch (roiaware_pool3d.cpp:125-135,
 * points_in_boxes_cuda.cu:79-105): out int32 [B,M,T] multi-hot, every element written. */
int occb200_points_in_boxes_batch(const float *boxes, const float *pts, const float *trig, int32_t *out,
                                  int B, int T, int M, int contract, void *stream);

/* HOST helper: trig[i] = {cosf(a), sinf(a)}, a = float(double(rz_i) + M_PI/2), with the host libm,
 * exactly as points_in_boxes_cpu.cpp:16-23 evaluates it.  boxes7 / trig are HOST pointers. */
void occb200_host_box_trig(const float *boxes7, int64_t n, float *trig);

/* ---- A7 / A8: Voxelization ------------------------------------------------------------ */

/*
 * Replaces voxel_layer.dynamic_voxelize (mmdet3d/ops/voxel/src/voxelization.cpp:8,
 * voxelization.h:71-83, voxelization_cuda.cu:24-65,332-375).
 * points [N, num_features] (dtype 0 = f32, 1 = f64), only columns 0..2 are read;
 * voxel_size[3], coors_range[6] are HOST arrays; coors int32 [N,3] in z,y,x order,
 * out-of-range points CLAMPED into the edge voxels (this fork's behaviour).
 */
int occb200_dynamic_voxelize(const void *points, int dtype, int64_t N, int num_features,
                             const float *voxel_size, const float *coors_range, int32_t *coors,
                             void *stream);

/*
 * Replaces voxel_layer.hard_voxelize (voxelization.cpp:7, voxelization.h:51-69,
 * voxelization_cuda.cu:67-184,188-330; CPU semantics voxelization_cpu.cpp:43-142).
 * points f32 [N,C]; caller allocates and ZERO-fills voxels f32 [max_voxels,max_points,C],
 * coors int32 [max_voxels,3], num_points_per_voxel int32 [max_voxels] (voxelize.py:46-50).
 * workspace: device scratch of occb200_hard_voxelize_workspace_bytes(N) bytes.
 * *voxel_num_host (HOST int) receives the voxel count; this call synchronises `stream`
 * (the reference returns the count, too).
 */
int64_t occb200_hard_voxelize_workspace_bytes(int64_t N);
int occb200_hard_voxelize(const float *points, int64_t N, int C, const float *voxel_size,
                          const float *coors_range, int max_points, int max_voxels, float *voxels,
                          int32_t *coors, int32_t *num_points_per_voxel, void *workspace,
                          int64_t workspace_bytes, int *voxel_num_host, void *stream);

/* ---- A6 / A9: unique voxels + scatter reductions -------------------------------------- */

/*
 * Sorted-unique of integer coordinate rows (at::unique_dim(..., sorted, inverse, counts) in
 * scatter_points_cuda.cu:204-205; torch.unique(coors, dim=0, ...) in ops/sst/sst_ops.py:156-158).
 * coors: int32 (coor_dtype 0) or int64 (coor_dtype 1) [N,K], 1 <= K <= 8.
 * mode 0 (scatter_v2): plain lexicographic unique.
 * mode 1 (DynamicScatter): rows with any negative coordinate are invalid (inverse -1); the first
 *         unique row is then dropped unconditionally -- the invalid group when there is one, else
 *         the lexicographically smallest valid voxel (scatter_points_cuda.cu:202-210).
 * mode 2 (DynamicScatter with 4-column batched coords, scatter_points.py:83-99): as mode 1 applied
 *         per value of column 0 (batch index in [0, coors[N-1,0]]); output rows keep column 0.
 * Outputs, caller-allocated for N rows: uniq [<=N,K] (dtype of coors), inverse int32 [N] (-1 = no
 * voxel), counts int32 [<=N], and the reduction plan: order int32 [N] (point indices grouped by
 * voxel, ascending inside a voxel) and gstart int32 [<=N] (first position of voxel m in order).
 * *m_host (HOST) receives the number of unique rows; the call synchronises `stream` (output
 * shapes depend on it, exactly as in the reference).  The packed sort key must fit 63 bits.
 */
int64_t occb200_unique_workspace_bytes(int64_t N, int K);
int occb200_unique_rows(const void *coors, int coor_dtype, int64_t N, int K, int mode, void *uniq,
                        int32_t *inverse, int32_t *counts, int32_t *order, int32_t *gstart,
                        void *workspace, int64_t workspace_bytes, int64_t *m_host, void *stream);

/* The same with caller-provided bounds (modes 1 / 2): col_max HOST int64 [K] = the largest value a valid row may hold
 * in each column (DynamicScatter knows its grid: round((range[3:] - range[:3]) / voxel_size) - 1, scatter_points.py:53-61;
 * the batch column of mode 2 is bounded by the batch size instead).  The key widths then need no min/max pass over the
 * rows and, in mode 1, no host round trip before the sort.  A row beyond a bound is detected on the device and the
 * call falls back to the unbounded path: the result is always that of occb200_unique_rows.  col_max NULL = unbounded. */
int occb200_unique_rows_bounded(const void *coors, int coor_dtype, int64_t N, int K, int mode, const int64_t *col_max,
                                void *uniq, int32_t *inverse, int32_t *counts, int32_t *order, int32_t *gstart,
                                void *workspace, int64_t workspace_bytes, int64_t *m_host, void *stream);

/* Reduction plan (order, gstart, counts) from an existing inverse map int32 [N] with values in
 * [-1, M): what scatter_v2 needs when the caller passes unq_inv (sst_ops.py:159-160). No sync. */
int64_t occb200_plan_workspace_bytes(int64_t N);
int occb200_plan_from_inverse(const int32_t *inverse, int64_t N, int64_t M, int32_t *order,
                              int32_t *gstart, int32_t *counts, void *workspace,
                              int64_t workspace_bytes, void *stream);

/*
 * Segmented reduction of feats f32 [N,C] into out f32 [M,C] following a plan
 * (feats_reduce_kernel scatter_points_cuda.cu:80-103 + mean division :228-229;
 *  torch_scatter scatter / scatter_max in sst_ops.py:171-176).
 * Deterministic: each voxel's points are combined in ascending point index by one (sub-)warp.
 * reduce: OCCB200_SUM / MEAN / MAX; mean divides by max(count,1); max of an empty voxel is -inf.
 * argmax (optional, int32 [M,C]): smallest point index attaining the max (:135-160), N if none.
 */
int occb200_segment_reduce(const float *feats, int64_t N, int C, const int32_t *order,
                           const int32_t *gstart, const int32_t *counts, int64_t M, int reduce,
                           float *out, int32_t *argmax, void *stream);

/*
 * Replaces voxel_layer.dynamic_point_to_voxel_backward (voxelization.cpp:10,
 * scatter_points_cuda.cu:236-303).  grad_feats f32 [N,C] is fully written (zero where no
 * gradient flows).  MAX: the gradient goes to the smallest point index whose feature equals the
 * voxel max; pass the forward's argmax, or NULL to recompute it from feats / reduced_feats
 * (stream-ordered scratch allocation inside the call).
 */
int occb200_segment_reduce_backward(float *grad_feats, const float *grad_reduced, const float *feats,
                                    const float *reduced_feats, const int32_t *inverse,
                                    const int32_t *counts, const int32_t *argmax, int64_t N,
                                    int64_t M, int C, int reduce, void *stream);

/* The same two entry points for float64 features: the reference dispatches DynamicScatter over
 * AT_DISPATCH_FLOATING_TYPES (scatter_points_cuda.cu:215, 260, 276, 289). */
int occb200_segment_reduce_f64(const double *feats, int64_t N, int C, const int32_t *order,
                               const int32_t *gstart, const int32_t *counts, int64_t M, int reduce,
                               double *out, int32_t *argmax, void *stream);
int occb200_segment_reduce_backward_f64(double *grad_feats, const double *grad_reduced, const double *feats,
                                        const double *reduced_feats, const int32_t *inverse,
                                        const int32_t *counts, const int32_t *argmax, int64_t N,
                                        int64_t M, int C, int reduce, void *stream);

/* ---- A10: occ_ops --------------------------------------------------------------------- */

/* mmdet3d/ops/occ/occ_ops.py:53-93 quantize_points: rois f32 [R,roi_dim] (sizes in columns 4..6),
 * roi_idx int64 [N] with PyTorch index semantics (a negative index counts from the end); writes out_coor
 * int64 [N,3] (to_center == 0) or out_center f32 [N,3].  An index outside [-R, R) -- an IndexError in the
 * reference -- is never dereferenced: the row gets INT64_MIN / NaN and is counted in *n_bad (device, may be
 * NULL; the caller zeroes it). */
int occb200_quantize_points(const float *points, int64_t N, const float *rois, int64_t R, int roi_dim,
                            const int64_t *roi_idx, float voxel_size, const float *scale_wlh,
                            const float *offset_wlh, int to_center, int64_t *out_coor,
                            float *out_center, unsigned long long *n_bad, void *stream);

/* occ_ops.py:5-50 generate_dense_voxel_centers for R boxes: sizes f32 [R,3] (DEVICE),
 * dims int32 [R,3] and center_off int64 [R+1] (DEVICE, computed by the caller with
 * ceil(size*scale+offset / vs) in f32), centers f32 [center_off[R],3], ij-meshgrid order. */
int occb200_dense_voxel_centers(const float *sizes, const int32_t *dims, const int64_t *center_off,
                                int R, int64_t total, float voxel_size, const float *scale_wlh,
                                const float *offset_wlh, float *centers, void *stream);

/* The dense observation grids of sample_observation (mmdet3d/models/roi_heads/bbox_heads/occ_ae_head.py:100-127) for
 * all R ROIs in one launch: coors int64 [N,3] (occb200_quantize_points), roi_idx int64 [N] (an index outside [0, R)
 * matches no ROI), dims int32 [R,3] and off int64 [R+1] as for occb200_dense_voxel_centers; labels int64 [off[R]]
 * (zero-filled by the caller) receives 1 at (c0*Y + c1)*Z + c2 of the point's ROI where 0 <= c < dims. */
int occb200_observed_labels(const int64_t *coors, const int64_t *roi_idx, int64_t N, const int32_t *dims,
                            const int64_t *off, int64_t R, int64_t *labels, void *stream);

/* MirrorOccLabel (mmdet3d/datasets/pipelines/occ_pinelines.py:82-126) on the label layout of
 * occb200_annotate_batch: out[label_off[t] + f] = labels[...] with every unknown (0) voxel replaced by the label
 * of its mirror image across the x mid-plane, read from the unmodified grid.  status may be NULL; tracklets with
 * status != 0 are skipped.  max_voxels = the largest X*Y*Z (sizes the grid).  out must not alias labels. */
int occb200_mirror_occ_label(const int32_t *labels, const int64_t *label_off, const int32_t *dims,
                             const int32_t *status, int32_t T, int64_t max_voxels, int32_t *out, void *stream);

/* ---- A2-A5: batched tracklet annotation (the "ray-cast") ------------------------------ */

/* One box pose per tracklet-frame: box7 plus the yaw trigonometry the reference evaluates on
 * the host side of its tensor code (see occb200_host_pose_pack).  64 bytes. */
typedef struct occb200_pose {
  float box[7];           /* x, y, z_bottom, x_size, y_size, z_size, yaw  (tools/ctrl/utils.py:42) */
  float cos_pib, sin_pib; /* cosf/sinf(float(yaw + pi/2)), host libm  (points_in_boxes_cpu.cpp:19-20) */
  float cos_m, sin_m;     /* torch f32 cos/sin(-yaw)                  (lidar_box3d.py:163-164)       */
  float cos_p, sin_p;     /* torch f32 cos/sin(+yaw)                  (occ_annotate.py:490-491)      */
  float pad[3];
} occb200_pose_t;

/* One LiDAR of one sensor frame.  80 bytes. */
typedef struct occb200_sensor {
  int64_t ri_off;    /* offset (floats) of the H x W f32 range image in ri_pool (0 = no return)     */
  int64_t incl_off;  /* offset (floats) of the FLIPPED inclination table in incl_pool (occ_annotate.py:528) */
  int32_t H, W;
  float v2l[12];     /* rows 0..2 of inv(extrinsic) evaluated in f32 on the host (occ_annotate.py:158-160) */
  float azc;         /* f32 atan2(E[1,0], E[0,0]) evaluated on the host (occ_annotate.py:175)        */
  int32_t incl_mono; /* +1 table ascending, -1 descending, 0 unknown (linear scan)                   */
} occb200_sensor_t;

typedef struct occb200_annotate_args {
  int32_t T;                      /* tracklets                                                */
  int32_t L;                      /* LiDARs per sensor frame, reference order (occ_annotate.py:235) */
  int64_t F;                      /* tracklet-frames = trk_frame_off[T]                       */
  const int64_t *trk_frame_off;   /* [T+1]                                                    */
  const occb200_pose_t *poses;    /* [F]                                                      */
  const int32_t *frame_sf;        /* [F] sensor-frame index of each tracklet-frame            */
  const float *points;            /* [P, point_stride] candidate returns, ego frame of their frame */
  int32_t point_stride;           /* floats per point (>= 3; 6 for KITTI-format .bin rows)    */
  int32_t pad0;
  const int64_t *frame_pt_off;    /* [F+1] point range of each tracklet-frame                 */
  const occb200_sensor_t *sensors;/* [SF, L]                                                  */
  int64_t SF;                     /* sensor frames in `sensors`                               */
  const float *incl_pool;
  int64_t incl_len;               /* floats in incl_pool                                      */
  const float *ri_pool;
  int64_t pyr_tiles;              /* sum of occb200_pyramid_tiles(H, W) over the SF*L sensors; 0 = no pair culling */
  int64_t items_cap;              /* reserved (ABI v5: bound of the work list; v6 sizes it from `bricks`)   */
  double voxel_size;              /* python float of --voxel-size (occ_annotate.py:215)       */
  const int64_t *label_off;       /* [T+1] slot of each tracklet in labels; slot size >= prod(ceil(max_frames(size)/vs)) */
  /* outputs */
  int32_t *labels;                /* [label_off[T]] 0 unknown / 1 occupied / 2 free, (x*Y+y)*Z+z inside the slot */
  int32_t *dims;                  /* [T,3]  X,Y,Z                                             */
  float *sizes;                   /* [T,3]  max box size over frames that have in-box points  */
  int32_t *status;                /* [T]    OCCB200_OK ...                                    */
  int64_t *n_unknown;             /* [T]    voxels tested for visibility (U)                  */
  int64_t *n_steps;               /* [T]    visibility tests actually evaluated (<= U*B*L; early exit) ; may be NULL */
  void *workspace;
  int64_t workspace_bytes;        /* >= occb200_annotate_workspace_bytes(T, F, label_off[T], SF, L, incl_len, pyr_tiles, bricks, max_pairs) */
  int32_t flags;                  /* bit 0: every visibility test in exact f64 (no f32 fast path);
                                     bit 1: no (frame, LiDAR) pair culling;
                                     bit 3: (tests) 64-entry recheck queue: overflowing tests are decided in place;
                                     bit 4: no brick-level culling (A/B measurement, parity);
                                     bit 5: scalar divisions as torch-CUDA evaluates them (x * (1/vs), see
                                            DESIGN.md section 4: "which arithmetic") instead of the CPU's x / vs */
  int32_t pad1;
  int64_t max_label_slots;        /* max_t (label_off[t+1] - label_off[t]), from the host copy of label_off; sizes the
                                     shared-memory bitsets.  0 = unknown (the 32 KB maximum is requested)        */
  /* ---- ABI v6: what the host already knows is passed in instead of being re-derived by small kernels ---- */
  uint8_t *labels_u8;             /* optional second label output, one byte per voxel, same layout as `labels`
                                     (the int32 of occ_annotate.py:581 is only needed at the file boundary);
                                     `labels` may be NULL when this is given                                */
  const int32_t *frame_trk;       /* [F] tracklet of each tracklet-frame                                    */
  const int64_t *pyr_off;         /* [SF*L + 1] prefix sum of occb200_pyramid_tiles(H, W) over the sensors
                                     (pyr_off[SF*L] == pyr_tiles); required when pyr_tiles > 0              */
  const int64_t *table_off;       /* [n_tables] incl_off of every DISTINCT inclination table (the frames of a
                                     segment share one table per LiDAR: occ_annotate.py:526-528)            */
  const int32_t *table_H;         /* [n_tables] its length                                                  */
  int32_t n_tables;
  int32_t max_pairs;              /* max_t (frames of t) * L (<= 4096): sizes the per-brick pair masks and the
                                     work lists                                                             */
  const int64_t *brick_off;       /* [T+1] prefix sum of occb200_grid_bricks() over the tracklets' label slots
                                     (the grid upper bound the slot was sized with): 4x4x4-voxel bricks are
                                     the unit of work and of culling in the ray-cast                        */
  int64_t bricks;                 /* brick_off[T], from the host copy                                        */
  int64_t n_points;               /* frame_pt_off[F], from the host copy: the crop kernel's grid             */
  /* ---- ABI v7 ---- */
  const uint8_t *ri_tile_live;    /* optional [pyr_tiles]: one byte per 8x32-pixel pyramid tile (image e at
                                     pyr_off[e], tile (tr, tc) at tr * ceil(W/32) + tc), 0 = no visibility test of
                                     the batch can read a pixel of the tile (as marked by occb200_pull_windows with
                                     the windowed upload): the max-pyramid skips those tiles instead of reading
                                     the whole 2.6 MB per frame the reference loads (occ_annotate.py:502-533).
                                     NULL = every tile is read                                               */
} occb200_annotate_args_t;

int64_t occb200_annotate_workspace_bytes(int32_t T, int64_t F, int64_t total_label_slots, int64_t SF,
                                         int32_t L, int64_t incl_len, int64_t pyr_tiles, int64_t bricks,
                                         int32_t max_pairs);
/* HOST helper: 4x4x4-voxel bricks of a grid of X x Y x Z voxels (for args.brick_off). */
int64_t occb200_grid_bricks(int32_t X, int32_t Y, int32_t Z);
/* HOST helper: tiles the max-pyramid of one H x W range image needs (for args.pyr_tiles). */
int64_t occb200_pyramid_tiles(int32_t H, int32_t W);

/*
 * What tools/occ/occ_annotate.py does per tracklet (get_local_point_list :91-138 and
 * OccAnnotator.annotate_trk :344-568), for T tracklets in one call: crop the candidate points
 * to the frame's box, move them to the box frame, voxelise, and label every voxel without a
 * point free/unknown by the reference's range-image visibility test over all frames and LiDARs.
 * No host synchronisation.  total_label_slots (= label_off[T]) is passed by value so that the
 * call needs no device->host read.
 */
int occb200_annotate_batch(const occb200_annotate_args_t *args, int64_t total_label_slots, void *stream);

/* After occb200_annotate_batch with the same args: out_host[0] = visibility tests the f32 fast path could
 * not decide and handed to the exact f64 recheck, out_host[1] = capacity of that queue (tests beyond it
 * are decided in place).  Synchronises `stream`. */
int occb200_annotate_queue_stats(const occb200_annotate_args_t *args, int64_t total_label_slots,
                                 int64_t *out_host, void *stream);

/* --save-mean-var support (tools/occ/occ_annotate.py:627-645).  Call after occb200_annotate_batch with the SAME
 * args (it reads the final grids from the workspace and args->status).  For each of the N candidate points:
 * loc_out f32 [N,3] = box-frame coordinates (occ_annotate.py:117-122), q_out int32 [N,4] = (tracklet, qx, qy, qz)
 * raw quantised coordinates (:425) of a point the reference keeps (:430-431), (-1,0,0,0) otherwise. */
int occb200_annotate_point_voxels(const occb200_annotate_args_t *args, int64_t total_label_slots, float *loc_out,
                                  int32_t *q_out, void *stream);

/*
 * Range-image builder: waymo_open_dataset.utils.range_image_utils.build_range_image_from_point_cloud (v1.2.0) as
 * the reference's converter calls it for *_RANGE_IMAGE_MERGE_VIRTUAL (tools/data_converter/waymo_converter.py:632-670).
 * One descriptor per image; the caller concatenates the two returns' points (:641-652) on its side.
 */
typedef struct {
  double v2l[12];     /* rows 0..2 of inv(extrinsic), evaluated in f64 on the host                          */
  double azc;         /* atan2(E[1,0], E[0,0]) in f64                                                         */
  int64_t incl_off;   /* offset of the image's inclination table in incl_pool, already reversed (:659)       */
  int64_t ri_off;     /* offset of the image in ri_pool (float index)                                         */
  int32_t H, W;
  int32_t mono;       /* -1 table strictly descending, +1 strictly ascending, 0 unknown (linear argmin)       */
  int32_t pad;
} occb200_ri_desc_t;

/* points f32 [*, point_stride] vehicle frame, image b owns rows [pt_off[b], pt_off[b+1]); max_points = the
 * largest of those counts (sizes the grid).  ri_pool f32 [ri_len] receives, per image, min range per pixel and
 * 0 where no point lands.  *n_bad (device) = points whose column fell outside [0, W) -- TF raises there; they
 * are skipped.  Asynchronous on `stream`. */
int occb200_build_range_images(const float *points, int point_stride, const int64_t *pt_off, int64_t max_points,
                               const occb200_ri_desc_t *desc, int32_t n_images, const float *incl_pool,
                               float *ri_pool, int64_t ri_len, unsigned long long *n_bad, void *stream);

/* ---- candidate selection from whole-frame clouds (occ_annotate.py:96-112 reads the full frame per tracklet-frame) ---- */

/* One candidate sphere: a box alive in a frame.  32 bytes. */
typedef struct occb200_cand_box {
  float cx, cy, cz;   /* sphere centre (box centre), ego frame of the frame                               */
  float r2;           /* squared radius: must contain the box (half diagonal + margin)                    */
  int32_t tf;         /* tracklet-frame this sphere feeds (informational; the output order is [frame][box]) */
  int32_t pad[3];
} occb200_cand_box_t;

/* Points per CTA chunk / warps per CTA of the selection kernel: the caller sizes `counts` with them. */
int occb200_candidate_chunk(void);
int occb200_candidate_warps(void);

/*
 * clouds f32 [*, stride]: the NF frame clouds of a segment back to back, frame f = rows [cloud_off[f], cloud_off[f+1]);
 * boxes: the spheres of frame f = boxes[frame_box_off[f] .. frame_box_off[f+1]).  With nchunk_f =
 * ceil(cloud size / occb200_candidate_chunk()) and W = occb200_candidate_warps(), the count of sphere k of frame f,
 * chunk c, warp w lives at counts[cnt_off[f] + ((k * nchunk_f + c) * W + w)].
 * Pass 1 (out_points == NULL) fills `counts`.  The caller takes the exclusive prefix sum over the whole array
 * (int64, same indexing) and calls again with out_points: every sphere's hits are then written to
 * out_points[prefix .. ) in cloud order, out_stride floats per point -- i.e. the candidate lists of the
 * tracklet-frames, contiguous, in [frame][box] order.  max_cloud = the largest cloud (sizes the grid).
 */
int occb200_select_candidates(const float *clouds, int stride, const int64_t *cloud_off, int32_t NF,
                              int64_t max_cloud, const occb200_cand_box_t *boxes,
                              const int64_t *frame_box_off, const int64_t *cnt_off, int32_t *counts,
                              const int64_t *prefix, float *out_points, int out_stride, void *stream);

/* ---- windowed upload of the range images (occ_annotate.py:502-533 loads every whole image; the visibility test
 *      reads a small window of each) ---------------------------------------------------------------------------- */

/* HOST helper.  mask8 (ceil(ri_len / 8) bytes, zeroed by the caller) receives 1 for every 32-byte block -- floats
 * [8k, 8k + 8) of ri_pool -- that the visibility test (occ_annotate.py:141-201, 541-547) of any voxel of any tracklet
 * of the batch can read: per (tracklet-frame, LiDAR, sub-box of the tracklet's centre box) the pixel footprint
 * bracketed from the sub-box's 8 corners.  trk_smax f32 [T,3] = the max box size over all frames of each tracklet;
 * sub_edge = sub-box edge in metres (<= 0: 0.8 m, what occb200_pull_windows uses on the device; 1e9: one box per
 * tracklet-frame, 12 % more bytes for 1/40 of the work).  All pointers are HOST arrays in the layout of
 * occb200_annotate_args_t.  Uploading only the marked blocks (occb200_host_mask_to_blocks + occb200_host_gather_blocks
 * + occb200_scatter_blocks) into a zero-initialised ri_pool gives the same labels, dims and statuses as uploading every
 * image; n_steps may differ (the culls see zeros outside the windows). */
int occb200_host_window_mark(int32_t T, int32_t L, const int64_t *trk_frame_off, const occb200_pose_t *poses,
                             const int32_t *frame_sf, const occb200_sensor_t *sensors, int64_t SF,
                             const float *incl_pool, const float *trk_smax, double voxel_size, int64_t ri_len,
                             uint8_t *mask8, float sub_edge);

/* HOST helper.  The ascending list of 16-float blocks (64 bytes) that hold a marked 8-float block; returns their
 * number.  out has room for ceil(n8 / 2) entries. */
int64_t occb200_host_mask_to_blocks(const uint8_t *mask8, int64_t n8, uint32_t *out);

/* HOST helper.  staging[16 i .. 16 i + 16) = block block_idx[i] of the pool, read from the source arrays: part j
 * holds floats [part_off[j], part_off[j] + part_len[j]) of the pool at part_ptr[j].  block_idx ascending;
 * part_off ascending multiples of 16; floats past a part's end are zero-filled. */
int occb200_host_gather_blocks(const uint32_t *block_idx, int64_t n_blocks, const int64_t *part_off,
                               const int64_t *part_len, const float *const *part_ptr, int32_t n_parts,
                               float *staging);

/* HOST.  dst[dst_off[i] ..) = src[i][0 .. bytes[i]) for n parts, copied in parallel (OpenMP): the one-shot API stages its
 * pageable inputs (the per-tracklet candidate point arrays the reference holds as separate tensors,
 * occ_annotate.py:96-138) in one pinned buffer with this before a single asynchronous H2D copy. */
int occb200_host_copy_parts(const void *const *src, const int64_t *bytes, const int64_t *dst_off, int32_t n, void *dst);

/* DEVICE.  ri_pool[16 block_idx[i] + j] = blocks[16 i + j], j < 16 (clipped at ri_len): scatters the uploaded
 * blocks to their place in the dense pool.  blocks and ri_pool 16-byte aligned.  Asynchronous on `stream`. */
int occb200_scatter_blocks(const float *blocks, const uint32_t *block_idx, int64_t n_blocks, float *ri_pool,
                           int64_t ri_len, void *stream);

/* The same upload with no host work.  occb200_window_mask_words(ri_len) uint32 of device scratch hold one bit per
 * 32-byte block (8 floats) of the pool.  occb200_pull_windows marks, ON THE DEVICE, the blocks the visibility test
 * of the batch can read -- per (tracklet-frame, LiDAR, sub-box of <= 0.8 m of the tracklet's centre box) the pixel
 * footprint bracketed from the sub-box's 8 corners -- and reads exactly those blocks from `ri_host`: the PINNED
 * HOST array holding the range images in the layout of ri_pool (a kernel may read pinned host memory under unified
 * addressing; the traffic crosses PCIe once, without a staging copy), storing them at the same offsets of ri_pool
 * (device, zero-initialised by the caller once).  `args` supplies the DEVICE metadata of occb200_annotate_batch
 * (T, L, F, SF, trk_frame_off, poses, frame_sf, sensors, incl_pool, voxel_size); trk_smax f32 [T,3] (device) = the
 * max box size over all frames of each tracklet.  ri_len % 8 == 0.  *pulled_blocks (device, optional, zeroed by
 * the caller) accumulates the 32-byte blocks read.  Asynchronous on `stream`. */
int64_t occb200_window_mask_words(int64_t ri_len);
int occb200_pull_windows(const occb200_annotate_args_t *args, const float *trk_smax, const float *ri_host,
                         float *ri_pool, int64_t ri_len, uint32_t *mask, unsigned long long *pulled_blocks,
                         uint8_t *tile_live, void *stream);
/* tile_live (device, optional, args->pyr_tiles bytes; needs args->pyr_off): set to 1 for every pyramid tile that holds
 * a marked block, 0 elsewhere -- what occb200_annotate_args_t.ri_tile_live takes for later calls on the same batch. */

/* ---- (f)3: dynamic_point_pool_mixed (mmdet3d/ops/dynamic_point_pool_op.py:63-113; the extension's kernel is not in
 *      the reference tree: PARITY UNPINNED, semantics from the extractor's own assertions, see csrc/point_pool.cu) ---- */

/* rois f32 [R,7]; ROI r scans the points perm[roi_lo[r] .. roi_hi[r]) (the caller sorts the points by batch index:
 * perm = that order, [lo, hi) = the run with the ROI's index); pts f32 [P,3]; extra_wlh HOST f32[3]; max_range = the
 * longest run (sizes the grid).  Pass 1 (out_feats == NULL): counts[r] += hits (caller-zeroed).  Pass 2: every hit is
 * written at base[r] + atomicAdd(cursor[r]) (cursor caller-zeroed): out_pidx / out_roi int64, out_feats f32 [.,13] =
 * xyz, ROI-local xyz, offsets to the six faces, is_in_margin.  Hits of a ROI arrive in arbitrary order. */
int occb200_point_pool(const float *rois, const int64_t *roi_lo, const int64_t *roi_hi, int64_t R, const float *pts,
                       const int64_t *perm, int64_t max_range, const float *extra_wlh, int32_t *counts,
                       const int64_t *base, int32_t *cursor, int64_t *out_pidx, int64_t *out_roi, float *out_feats,
                       void *stream);

/* HOST helper: fills poses[i] from boxes7 f32 [n,7] and torch-evaluated trig f32 [n,4]
 * (cos(-yaw), sin(-yaw), cos(yaw), sin(yaw)); cos_pib/sin_pib come from the host libm. */
void occb200_host_pose_pack(const float *boxes7, const float *trig4, int64_t n, occb200_pose_t *poses);

/*
 * The reference's point_cloud_to_range_image_idx (tools/occ/occ_annotate.py:141-201) as an
 * operator: points f64 [B,N,3]; v2l f32 [B,12] and azc f32 [B] host-evaluated as in
 * occb200_sensor_t; incl f32 [B,H] already flipped -> ri_idx int64 [B,N,2] (row, col; col may be
 * negative exactly as in the reference), ri_range f64 [B,N].
 */
int occb200_point_cloud_to_range_image_idx(const double *points, int B, int64_t N, const float *v2l,
                                           const float *azc, const float *incl, int H, int W,
                                           int64_t *ri_idx, double *ri_range, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OCC_B200_H_ */
