/*
 * occ_oracle.c -- CPU restatement of the reference's point -> occupancy hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (objectcentricocccompletion_b200) never does.
 *
 * Every function restates, scalar by scalar and in the reference's operation
 * order, what the reference (/root/reference, Ghostish/ObjectCentricOccCompletion)
 * computes when it runs on CPU tensors; the file:line it follows is cited on
 * each function.  Where the reference uses a torch matmul / einsum / norm, the
 * arithmetic is the FMA chain torch-CPU produces (re-checked
 * by oracle/validate_oracle.py against the real torch ops).  Build with
 * -ffp-contract=off so that nothing else is fused.
 *
 * Pinning: oracle/validate_oracle.py checks this file against (a) the
 * reference's own point_cloud_to_range_image_idx exec'd from source,
 * (b) the reference's compiled voxel_layer / points_in_boxes_cpu (oracle/_ref),
 * (c) the golden vectors of the reference's tests (tests/golden).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_SKIP_SHORT 1     /* len(trk) < 10                 occ_annotate.py:344 */
#define ORC_NO_POINTS 2      /* assert len(local_pc_list) > 0 occ_annotate.py:129 */
#define ORC_EMPTY_AFTER_FILTER 3 /* max() of empty tensor     occ_annotate.py:433 */
#define ORC_INDEX_ERROR 4    /* q < -dims -> IndexError       occ_annotate.py:436 */

/* ------------------------------------------------------------------------ */
/* A1  points_in_boxes                                                       */
/* ------------------------------------------------------------------------ */

/* mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp:16-41
 * (same arithmetic as points_in_boxes_cuda.cu:24-49).  cos/sin are the float
 * overloads; everything touching "2.0" or M_PI is promoted to double. */
static int check_pt_in_box3d(const float *pt, const float *box3d) {
  float x = pt[0], y = pt[1], z = pt[2];
  float cx = box3d[0], cy = box3d[1], cz = box3d[2];
  float w = box3d[3], l = box3d[4], h = box3d[5], rz = box3d[6];
  cz = (float)((double)cz + (double)h / 2.0);
  if ((double)fabsf(z - cz) > (double)h / 2.0) return 0;
  float rot_angle = (float)((double)rz + M_PI / 2);
  float cosa = cosf(rot_angle), sina = sinf(rot_angle);
  float shift_x = x - cx, shift_y = y - cy;
  float local_x = shift_x * cosa + shift_y * (-sina);
  float local_y = shift_x * sina + shift_y * cosa;
  int in_flag = ((double)local_x > -(double)l / 2.0) & ((double)local_x < (double)l / 2.0) &
                ((double)local_y > -(double)w / 2.0) & ((double)local_y < (double)w / 2.0);
  return in_flag;
}

/* points_in_boxes_cpu.cpp:43-69: out[T, M] = 0/1 */
void orc_points_in_boxes_cpu(int T, int M, const float *boxes, const float *pts, int32_t *out) {
  for (int i = 0; i < T; i++)
    for (int j = 0; j < M; j++) out[(long)i * M + j] = check_pt_in_box3d(pts + 3 * (long)j, boxes + 7 * i);
}

/* points_in_boxes_cuda.cu:51-77 + points_in_boxes.py:29-30: out[B, M] = first box hit, else -1 */
void orc_points_in_boxes_gpu(int B, int T, int M, const float *boxes, const float *pts, int32_t *out) {
  for (int b = 0; b < B; b++)
    for (int j = 0; j < M; j++) {
      int32_t idx = -1;
      for (int k = 0; k < T; k++)
        if (check_pt_in_box3d(pts + 3 * ((long)b * M + j), boxes + 7 * ((long)b * T + k))) {
          idx = k;
          break;
        }
      out[(long)b * M + j] = idx;
    }
}

/* points_in_boxes_cuda.cu:79-105 + points_in_boxes.py:109-110: out[B, M, T] multi-hot */
void orc_points_in_boxes_batch(int B, int T, int M, const float *boxes, const float *pts, int32_t *out) {
  for (int b = 0; b < B; b++)
    for (int j = 0; j < M; j++)
      for (int k = 0; k < T; k++)
        out[((long)b * M + j) * T + k] =
            check_pt_in_box3d(pts + 3 * ((long)b * M + j), boxes + 7 * ((long)b * T + k)) ? 1 : 0;
}

/* ------------------------------------------------------------------------ */
/* A7 / A8  Voxelization                                                     */
/* ------------------------------------------------------------------------ */

/* mmdet3d/ops/voxel/src/voxelization_cpu.cpp:7-41 (clamping fork; zyx order).
 * stride = number of feature columns of `points`. */
static void dynamic_voxelize_rows(const float *points, long n, int stride, const float *voxel_size,
                                  const float *coors_range, const int *grid_size, int32_t *coors) {
  for (long i = 0; i < n; i++) {
    for (int j = 0; j < 3; j++) {
      int c = (int)floorf((points[i * stride + j] - coors_range[j]) / voxel_size[j]);
      if (c < 0)
        c = 0;
      else if (c >= grid_size[j])
        c = grid_size[j] - 1;
      coors[i * 3 + (2 - j)] = c;
    }
  }
}

/* voxelization_cpu.cpp:144-169: grid = ceil((max-min)/vs) in float */
void orc_dynamic_voxelize(const float *points, long n, int stride, const float *voxel_size,
                          const float *coors_range, int32_t *coors) {
  int grid[3];
  for (int i = 0; i < 3; i++) grid[i] = (int)ceilf((coors_range[3 + i] - coors_range[i]) / voxel_size[i]);
  dynamic_voxelize_rows(points, n, stride, voxel_size, coors_range, grid, coors);
}

/* voxelization_cpu.cpp:43-142: grid = round(...); voxels in first-appearance order.
 * voxels [max_voxels, max_points, stride], coors [max_voxels,3], num [max_voxels]: caller zero-fills. */
int orc_hard_voxelize(const float *points, long n, int stride, const float *voxel_size,
                      const float *coors_range, int max_points, int max_voxels, float *voxels,
                      int32_t *coors, int32_t *num_points_per_voxel) {
  int grid[3];
  for (int i = 0; i < 3; i++) grid[i] = (int)roundf((coors_range[3 + i] - coors_range[i]) / voxel_size[i]);
  int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(n > 0 ? n : 1));
  dynamic_voxelize_rows(points, n, stride, voxel_size, coors_range, grid, tmp);
  size_t cells = (size_t)grid[0] * grid[1] * grid[2];
  int32_t *c2v = (int32_t *)malloc(sizeof(int32_t) * (cells > 0 ? cells : 1));
  for (size_t i = 0; i < cells; i++) c2v[i] = -1;
  int voxel_num = 0;
  for (long i = 0; i < n; i++) {
    const int32_t *c = tmp + 3 * i;
    if (c[0] == -1) continue;
    size_t cell = ((size_t)c[0] * grid[1] + c[1]) * grid[0] + c[2];
    int voxelidx = c2v[cell];
    if (voxelidx == -1) {
      voxelidx = voxel_num;
      if (max_voxels != -1 && voxel_num >= max_voxels) continue;
      voxel_num += 1;
      c2v[cell] = voxelidx;
      for (int k = 0; k < 3; k++) coors[voxelidx * 3 + k] = c[k];
    }
    int num = num_points_per_voxel[voxelidx];
    if (max_points == -1 || num < max_points) {
      for (int k = 0; k < stride; k++)
        voxels[((size_t)voxelidx * max_points + num) * stride + k] = points[i * stride + k];
      num_points_per_voxel[voxelidx] += 1;
    }
  }
  free(tmp);
  free(c2v);
  return voxel_num;
}

/* ------------------------------------------------------------------------ */
/* A6 / A9  unique rows + scatter reductions                                 */
/* ------------------------------------------------------------------------ */

typedef struct {
  const int64_t *rows;
  int k;
} rowcmp_ctx;
static rowcmp_ctx g_ctx; /* qsort has no context argument; oracle is single-threaded here */

static int rowcmp(const void *a, const void *b) {
  long ia = *(const long *)a, ib = *(const long *)b;
  const int64_t *ra = g_ctx.rows + ia * g_ctx.k, *rb = g_ctx.rows + ib * g_ctx.k;
  for (int j = 0; j < g_ctx.k; j++) {
    if (ra[j] < rb[j]) return -1;
    if (ra[j] > rb[j]) return 1;
  }
  return (ia > ib) - (ia < ib);
}

/* torch.unique(coors, dim=0, return_inverse=True, return_counts=True): rows sorted
 * lexicographically (sst_ops.py:156-158, scatter_points_cuda.cu:204-205).
 * Returns M; uniq [<=n, k], inv [n], cnt [<=n]. */
long orc_unique_rows(const int64_t *rows, long n, int k, int64_t *uniq, int64_t *inv, int64_t *cnt) {
  if (n == 0) return 0;
  long *order = (long *)malloc(sizeof(long) * n);
  for (long i = 0; i < n; i++) order[i] = i;
  g_ctx.rows = rows;
  g_ctx.k = k;
  qsort(order, n, sizeof(long), rowcmp);
  long m = 0;
  for (long t = 0; t < n; t++) {
    long i = order[t];
    int is_new = (t == 0);
    if (!is_new) {
      const int64_t *prev = rows + order[t - 1] * k;
      for (int j = 0; j < k; j++)
        if (prev[j] != rows[i * k + j]) {
          is_new = 1;
          break;
        }
    }
    if (is_new) {
      for (int j = 0; j < k; j++) uniq[m * k + j] = rows[i * k + j];
      cnt[m] = 0;
      m++;
    }
    inv[i] = m - 1;
    cnt[m - 1]++;
  }
  free(order);
  return m;
}

/* Segment reductions shared by scatter_v2 (sst_ops.py:171-176, torch_scatter
 * semantics: mean = sum / clamp(count, 1); max returns values) and DynamicScatter
 * (scatter_points_cuda.cu:80-103, 217-229: mean = sum / count; max init -inf).
 * map[i] == -1 skips the point.  mode: 0 sum, 1 mean, 2 max.  Accumulates in
 * point order in f32 (the reference's atomics have no defined order). */
void orc_segment_reduce(const float *feats, long n, int c, const int64_t *map, long m, int mode,
                        float *out, const int64_t *cnt) {
  for (long i = 0; i < m * c; i++) out[i] = (mode == 2) ? -INFINITY : 0.0f;
  for (long i = 0; i < n; i++) {
    long v = map[i];
    if (v < 0) continue;
    for (int j = 0; j < c; j++) {
      float f = feats[i * c + j];
      float *o = out + v * c + j;
      if (mode == 2)
        *o = fmaxf(*o, f); /* fmaxf ignores NaN like the CAS loop's `old < val` test */
      else
        *o += f;
    }
  }
  if (mode == 1)
    for (long v = 0; v < m; v++) {
      float d = (float)(cnt[v] < 1 ? 1 : cnt[v]);
      for (int j = 0; j < c; j++) out[v * c + j] /= d;
    }
}

/* DynamicScatter forward, scatter_points_cuda.cu:183-234:
 *  rows with any negative coord -> (-1,-1,-1); unique sorted; FIRST unique row is dropped
 *  unconditionally (:207-210); map = inv - 1.  coors int32 [n,3].
 * Returns M (after the drop).  voxel_coors [<=n,3] int32, map int32 [n], cnt int32 [<=n]. */
long orc_dynamic_scatter_fwd(const float *feats, const int32_t *coors, long n, int c, int mode,
                             float *voxel_feats, int32_t *voxel_coors, int32_t *map, int32_t *cnt) {
  if (n == 0) return 0;
  int64_t *clean = (int64_t *)malloc(sizeof(int64_t) * 3 * n);
  for (long i = 0; i < n; i++) {
    int neg = coors[i * 3] < 0 || coors[i * 3 + 1] < 0 || coors[i * 3 + 2] < 0;
    for (int j = 0; j < 3; j++) clean[i * 3 + j] = neg ? -1 : coors[i * 3 + j];
  }
  int64_t *uniq = (int64_t *)malloc(sizeof(int64_t) * 3 * n);
  int64_t *inv = (int64_t *)malloc(sizeof(int64_t) * n);
  int64_t *cn = (int64_t *)malloc(sizeof(int64_t) * n);
  long m_all = orc_unique_rows(clean, n, 3, uniq, inv, cn);
  long m = m_all - 1;
  for (long v = 0; v < m; v++) {
    for (int j = 0; j < 3; j++) voxel_coors[v * 3 + j] = (int32_t)uniq[(v + 1) * 3 + j];
    cnt[v] = (int32_t)cn[v + 1];
  }
  for (long i = 0; i < n; i++) {
    inv[i] -= 1;
    map[i] = (int32_t)inv[i];
  }
  orc_segment_reduce(feats, n, c, inv, m, mode, voxel_feats, cn + 1);
  free(clean);
  free(uniq);
  free(inv);
  free(cn);
  return m;
}

/* DynamicScatter backward, scatter_points_cuda.cu:236-303.  grad_feats zero-filled here.
 * max: gradient goes to the smallest point index whose feature equals the max (:135-179). */
void orc_dynamic_scatter_bwd(float *grad_feats, const float *grad_voxel, const float *feats,
                             const float *voxel_feats, const int32_t *map, const int32_t *cnt, long n,
                             long m, int c, int mode) {
  memset(grad_feats, 0, sizeof(float) * (size_t)n * c);
  if (n == 0 || m == 0) return;
  if (mode == 0 || mode == 1) {
    for (long i = 0; i < n; i++) {
      long v = map[i];
      if (v < 0) continue;
      for (int j = 0; j < c; j++)
        grad_feats[i * c + j] = (mode == 0) ? grad_voxel[v * c + j] : grad_voxel[v * c + j] / (float)cnt[v];
    }
  } else {
    int32_t *from = (int32_t *)malloc(sizeof(int32_t) * (size_t)m * c);
    for (long i = 0; i < m * c; i++) from[i] = (int32_t)n;
    for (long i = 0; i < n; i++) {
      long v = map[i];
      if (v < 0) continue;
      for (int j = 0; j < c; j++)
        if (feats[i * c + j] == voxel_feats[v * c + j] && (int32_t)i < from[v * c + j]) from[v * c + j] = (int32_t)i;
    }
    for (long v = 0; v < m; v++)
      for (int j = 0; j < c; j++)
        if (from[v * c + j] < n) grad_feats[(long)from[v * c + j] * c + j] = grad_voxel[v * c + j];
    free(from);
  }
}

/* ------------------------------------------------------------------------ */
/* A10  occ_ops.quantize_points / generate_dense_voxel_centers               */
/* ------------------------------------------------------------------------ */

/* mmdet3d/ops/occ/occ_ops.py:53-93.  rois [R, roi_dim] (batch_idx,x,y,z,w,l,h,ry,...);
 * out_coor int64 [n,3] or, when to_center, out_center f32 [n,3]. */
void orc_quantize_points(const float *points, long n, const float *rois, int roi_dim,
                         const int64_t *roi_idx, float voxel_size, const float *scale_wlh,
                         const float *offset_wlh, int to_center, int64_t *out_coor, float *out_center) {
  for (long i = 0; i < n; i++) {
    const float *roi = rois + roi_idx[i] * roi_dim;
    for (int j = 0; j < 3; j++) {
      float size = roi[4 + j] * scale_wlh[j] + offset_wlh[j];
      float mn = -size / 2.0f;
      float q = floorf((points[i * 3 + j] - mn) / voxel_size);
      int64_t qi = (int64_t)q;
      if (to_center)
        out_center[i * 3 + j] = (float)qi * voxel_size + mn + voxel_size / 2.0f;
      else
        out_coor[i * 3 + j] = qi;
    }
  }
}

/* occ_ops.py:5-50 for one bbox size: dims = ceil(size/vs); centres in ij-meshgrid order.
 * Returns the number of centres; centers f32 [X*Y*Z, 3] (may be NULL to query dims). */
long orc_dense_voxel_centers(const float *bbox_size, float voxel_size, const float *scale_wlh,
                             const float *offset_wlh, int *dims, float *centers) {
  float size[3];
  for (int j = 0; j < 3; j++) {
    size[j] = bbox_size[j] * scale_wlh[j] + offset_wlh[j];
    dims[j] = (int)ceilf(size[j] / voxel_size);
  }
  long nvox = (long)dims[0] * dims[1] * dims[2];
  if (!centers) return nvox;
  long t = 0;
  for (int x = 0; x < dims[0]; x++)
    for (int y = 0; y < dims[1]; y++)
      for (int z = 0; z < dims[2]; z++, t++) {
        int q[3] = {x, y, z};
        for (int j = 0; j < 3; j++) {
          float mn = -size[j] / 2.0f;
          centers[t * 3 + j] = (float)q[j] * voxel_size + mn + voxel_size / 2.0f;
        }
      }
  return nvox;
}

/* ------------------------------------------------------------------------ */
/* A5  point_cloud_to_range_image_idx (single frame, f64)                    */
/* ------------------------------------------------------------------------ */

/* tools/occ/occ_annotate.py:141-201 for one point.
 *  v2l: f32 3x4 rows of inv(extrinsic) (the inverse itself is a host torch-CPU op, :158-160)
 *  azc: f32 atan2(E[1,0], E[0,0]) (host torch-CPU op, :175)
 *  incl: H inclinations exactly as handed to the function (the caller flips, :528). */
static void project_point(const double ego[3], const float *v2l, float azc, const float *incl, int H,
                          int W, int64_t *row, int64_t *col, double *range) {
  double p[3];
  for (int k = 0; k < 3; k++) { /* einsum 'bij,bkj->bik' == FMA chain over j, then + translation (:164) */
    double acc = ego[0] * (double)v2l[k * 4 + 0];
    acc = fma(ego[1], (double)v2l[k * 4 + 1], acc);
    acc = fma(ego[2], (double)v2l[k * 4 + 2], acc);
    p[k] = acc + (double)v2l[k * 4 + 3];
  }
  double xy_norm = sqrt(fma(p[1], p[1], p[0] * p[0])); /* :165 */
  double inc = atan2(p[2], xy_norm);                     /* :166 */
  int best = 0;
  double bestd = INFINITY;
  for (int h = 0; h < H; h++) { /* :168-173 argmin, first index wins ties */
    double d = fabs(inc - (double)incl[h]);
    if (d < bestd) {
      bestd = d;
      best = h;
    }
  }
  double az = atan2(p[1], p[0]) + (double)azc; /* :176-178 */
  const double two_pi_f32 = (double)(2.0f * (float)M_PI); /* mask.to(float32) * 2 * np.pi stays float32 (:182,:185) */
  int gt = az > M_PI, lt = az < -M_PI;
  if (gt) az = az - two_pi_f32;
  if (lt) az = az + two_pi_f32;
  double colf = (double)W - 1.0 + 0.5 - (az + M_PI) / (2.0 * M_PI) * (double)W; /* :187-189 */
  colf = nearbyint(colf);           /* torch.round: half to even (:190) */
  colf = fmod(colf, (double)W);     /* :191 */
  *row = best;
  *col = (int64_t)(int32_t)colf;    /* .to(torch.int32) (:191-193) */
  *range = sqrt(fma(p[2], p[2], fma(p[1], p[1], p[0] * p[0]))); /* :198 */
}

/* Batched wrapper with the reference's signature: points f64 [B,N,3], v2l f32 [B,12],
 * azc f32 [B], incl f32 [B,H] -> ri_idx int64 [B,N,2], ri_range f64 [B,N]. */
void orc_point_cloud_to_range_image_idx(const double *points, int B, long N, const float *v2l,
                                        const float *azc, const float *incl, int H, int W,
                                        int64_t *ri_idx, double *ri_range) {
  for (int b = 0; b < B; b++)
    for (long i = 0; i < N; i++) {
      long t = (long)b * N + i;
      project_point(points + 3 * t, v2l + 12 * b, azc[b], incl + (long)H * b, H, W, ri_idx + 2 * t,
                    ri_idx + 2 * t + 1, ri_range + t);
    }
}

/* ------------------------------------------------------------------------ */
/* A2-A5  annotate one tracklet                                              */
/* ------------------------------------------------------------------------ */

typedef struct {
  int64_t ri_off;   /* offset (floats) of the H x W range image in ri_pool          */
  int64_t incl_off; /* offset (floats) of the FLIPPED inclination table in incl_pool */
  int32_t H, W;
  float v2l[12];    /* f32 inverse extrinsic, rows 0..2                              */
  float azc;        /* f32 atan2(E[1,0], E[0,0])                                     */
  int32_t pad;
} orc_sensor_t;

/*
 * tools/occ/occ_annotate.py: get_local_point_list :91-138 and annotate_trk :344-568.
 *
 * boxes [B,7] f32; trig [B,4] f32 = cos(-rz), sin(-rz), cos(rz), sin(rz) as computed by
 * torch-CPU float32 (lidar_box3d.py:163-164, occ_annotate.py:490-491 -- host ops);
 * pts f32 [P,3] with pt_off [B+1]; sensors [B*L] in the reference's LiDAR order (:235).
 * labels int32 [cap]; dims/size outputs.  loc_out / keep_out (optional, [P,3] / [P]) expose
 * the local points and the in-box flag for intermediate checks.
 */
int orc_annotate_tracklet(int B, const float *boxes, const float *trig, const float *pts,
                          const int64_t *pt_off, const orc_sensor_t *sensors, int L,
                          const float *incl_pool, const float *ri_pool, double voxel_size,
                          int32_t *dims_out, float *size_out, int32_t *labels, long cap,
                          int64_t *n_unknown_out, float *loc_out, int8_t *keep_out) {
  dims_out[0] = dims_out[1] = dims_out[2] = 0;
  *n_unknown_out = 0;
  if (B < 10) return ORC_SKIP_SHORT; /* :344 */
  long P = pt_off[B] - pt_off[0];
  float *loc = (float *)malloc(sizeof(float) * 3 * (size_t)(P > 0 ? P : 1));
  long nloc = 0;
  float size[3] = {-INFINITY, -INFINITY, -INFINITY};
  int kept_frames = 0;
  const float *pbase = pts + 3 * pt_off[0];
  for (int i = 0; i < B; i++) { /* :96-128 */
    const float *box = boxes + 7 * i;
    long n0 = pt_off[i] - pt_off[0], n1 = pt_off[i + 1] - pt_off[0];
    long before = nloc;
    float c = trig[i * 4 + 0], s = trig[i * 4 + 1];
    for (long j = n0; j < n1; j++) {
      const float *p = pbase + 3 * j;
      int in = check_pt_in_box3d(p, box); /* :109-110, inbox_inds == 0 with a single box */
      if (keep_out) keep_out[j] = (int8_t)in;
      if (!in) continue;
      float tx = p[0] + (-box[0]), ty = p[1] + (-box[1]), tz = p[2] + (-box[2]); /* :117-120 */
      /* points @ [[c,-s,0],[s,c,0],[0,0,1]] as the sgemm FMA chain (lidar_box3d.py:165-184) */
      float lx = fmaf(tz, 0.0f, fmaf(ty, s, tx * c));
      float ly = fmaf(tz, 0.0f, fmaf(ty, c, tx * (-s)));
      float lz = fmaf(tz, 1.0f, fmaf(ty, 0.0f, tx * 0.0f));
      loc[3 * nloc + 0] = lx;
      loc[3 * nloc + 1] = ly;
      loc[3 * nloc + 2] = lz;
      if (loc_out) {
        loc_out[3 * j + 0] = lx;
        loc_out[3 * j + 1] = ly;
        loc_out[3 * j + 2] = lz;
      }
      nloc++;
    }
    if (nloc == before) continue; /* :111-112 */
    kept_frames++;
    for (int k = 0; k < 3; k++) size[k] = fmaxf(size[k], box[3 + k]); /* :132-133 box_mode="max" */
  }
  if (kept_frames == 0) {
    free(loc);
    return ORC_NO_POINTS; /* :129, :356-359 */
  }
  float vsf = (float)voxel_size;
  int dims[3];
  float mb[3];
  for (int k = 0; k < 3; k++) {
    dims[k] = (int)ceilf(size[k] / vsf); /* :414-416 */
    size_out[k] = size[k];
    dims_out[k] = dims[k];
  }
  mb[0] = size[0] * -0.5f; /* corners with yaw 0, min over 8 (lidar_box3d.py:80-92, :422-423) */
  mb[1] = size[1] * -0.5f;
  mb[2] = size[2] * 0.0f;
  long V = (long)dims[0] * dims[1] * dims[2];
  if (V > cap) {
    free(loc);
    return -1;
  }
  uint8_t *occ = (uint8_t *)calloc((size_t)(V > 0 ? V : 1), 1);
  long nkept = 0;
  int status = ORC_OK;
  for (long j = 0; j < nloc; j++) { /* :425-438 */
    int64_t q[3];
    int ok = 1;
    for (int k = 0; k < 3; k++) {
      q[k] = (int64_t)floorf((loc[3 * j + k] - mb[k]) / vsf);
      if (!(q[k] < dims[k])) ok = 0; /* only the upper bound is tested (:430-431) */
    }
    if (!ok) continue;
    nkept++;
    for (int k = 0; k < 3; k++) {
      if (q[k] < 0) q[k] += dims[k]; /* PyTorch negative-index wrap (:436) */
      if (q[k] < 0) status = ORC_INDEX_ERROR;
    }
    if (status == ORC_INDEX_ERROR) break;
    occ[(q[0] * dims[1] + q[1]) * dims[2] + q[2]] = 1;
  }
  free(loc);
  if (status == ORC_OK && nkept == 0) status = ORC_EMPTY_AFTER_FILTER; /* :433-435 */
  if (status != ORC_OK) {
    free(occ);
    return status;
  }
  long U = 0;
  for (long f = 0; f < V; f++) U += !occ[f];
  *n_unknown_out = U;
  const double vs = voxel_size;
  for (long f = 0; f < V; f++) {
    if (occ[f]) {
      labels[f] = 1; /* :563 */
      continue;
    }
    long x = f / ((long)dims[1] * dims[2]), y = (f / dims[2]) % dims[1], z = f % dims[2];
    double cen[3]; /* :467-471, evaluated left to right in f64 */
    cen[0] = (double)x * vs + (double)mb[0] + vs / 2;
    cen[1] = (double)y * vs + (double)mb[1] + vs / 2;
    cen[2] = (double)z * vs + (double)mb[2] + vs / 2;
    int free_ = 0;
    for (int c = 0; c < L && !free_; c++)     /* :525 ; OR over LiDARs :552-556 */
      for (int i = 0; i < B && !free_; i++) { /* :479 ; OR over frames :550    */
        const float *box = boxes + 7 * i;
        double rc = (double)trig[i * 4 + 2], rs = (double)trig[i * 4 + 3]; /* :490-496 */
        double ego[3]; /* centres @ [[c,-s,0],[s,c,0],[0,0,1]] + origin (:497-498) */
        ego[0] = fma(cen[2], 0.0, fma(cen[1], rs, cen[0] * rc)) + (double)box[0];
        ego[1] = fma(cen[2], 0.0, fma(cen[1], rc, cen[0] * (-rs))) + (double)box[1];
        ego[2] = fma(cen[2], 1.0, fma(cen[1], 0.0, cen[0] * 0.0)) + (double)box[2];
        const orc_sensor_t *sn = sensors + (long)i * L + c;
        int64_t row, col;
        double rng;
        project_point(ego, sn->v2l, sn->azc, incl_pool + sn->incl_off, sn->H, sn->W, &row, &col, &rng);
        if (col < 0) col += sn->W; /* negative index wrap (:543) */
        double riv = (double)ri_pool[sn->ri_off + row * sn->W + col];
        if (riv >= rng) free_ = 1; /* :547 */
      }
    labels[f] = free_ ? 2 : 0; /* :558-562 */
  }
  free(occ);
  return ORC_OK;
}

/* Batch driver used by bench.py's CPU baseline: tracklets are independent
 * (occ_annotate.py:649-671 shards them over worker processes), so an OpenMP
 * loop over tracklets mirrors --workers N. */
typedef struct {
  int32_t B;
  int32_t pad;
  int64_t frame0;     /* first tracklet-frame in boxes/trig/pt_off  */
  int64_t label_off;  /* offset of the tracklet's slot in labels    */
  int64_t label_cap;
} orc_trk_t;

void orc_annotate_batch(int T, const orc_trk_t *trk, const float *boxes, const float *trig,
                        const float *pts, const int64_t *pt_off, const int32_t *frame_sf,
                        const orc_sensor_t *sensors, int L, const float *incl_pool,
                        const float *ri_pool, double voxel_size, int32_t *dims_out, float *size_out,
                        int32_t *labels, int32_t *status, int64_t *n_unknown, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < T; t++) {
    int B = trk[t].B;
    int64_t f0 = trk[t].frame0;
    orc_sensor_t *sn = (orc_sensor_t *)malloc(sizeof(orc_sensor_t) * (size_t)(B > 0 ? B : 1) * L);
    for (int i = 0; i < B; i++)
      for (int c = 0; c < L; c++) sn[(long)i * L + c] = sensors[(long)frame_sf[f0 + i] * L + c];
    status[t] = orc_annotate_tracklet(B, boxes + 7 * f0, trig + 4 * f0, pts, pt_off + f0, sn, L, incl_pool,
                                      ri_pool, voxel_size, dims_out + 3 * t, size_out + 3 * t,
                                      labels + trk[t].label_off, trk[t].label_cap, n_unknown + t, NULL, NULL);
    free(sn);
  }
}


/* ---------------------------------------------------------------------------------------------
 * Range-image builder: waymo_open_dataset.utils.range_image_utils.build_range_image_from_point_cloud
 * (waymo-open-dataset-tf-2-1-0 == 1.2.0, requirements/optional.txt; NOT vendored under the reference: this
 * restates its published algorithm, PARITY UNPINNED against the TF code itself), as called for one LiDAR of one
 * frame at tools/data_converter/waymo_converter.py:653-668.  f64 throughout after the casts; the products
 * accumulate as an FMA chain like the torch port of the same function (occ_annotate.py:161-164).
 * v2l = rows 0..2 of inv(f64(extrinsic)) (host), azc = atan2(E[1,0], E[0,0]) in f64, incl = table reversed (:659).
 * ri [H*W] = min range per pixel, 0 where no point lands; rows/cols/ranges per point (optional).
 * Returns the number of points whose column is outside [0, W) (TF asserts; they are skipped). */
long orc_build_range_image(const float *points, long n, int stride, const double *v2l, double azc,
                           const float *incl, int H, int W, float *ri, int32_t *rows, int32_t *cols,
                           float *ranges) {
  const double two_pi = 6.28318530717958647692;
  long bad = 0;
  for (long k = 0; k < (long)H * W; k++) ri[k] = INFINITY;
  for (long i = 0; i < n; i++) {
    double x = points[i * stride], y = points[i * stride + 1], z = points[i * stride + 2], p[3];
    for (int r = 0; r < 3; r++)
      p[r] = fma(z, v2l[4 * r + 2], fma(y, v2l[4 * r + 1], x * v2l[4 * r])) + v2l[4 * r + 3];
    double inc = atan2(p[2], sqrt(fma(p[1], p[1], p[0] * p[0])));
    int row = 0;
    double best = INFINITY;
    for (int h = 0; h < H; h++) {
      double d = fabs(inc - (double)incl[h]);
      if (d < best) { best = d; row = h; }
    }
    double az = atan2(p[1], p[0]) + azc;
    int gt = az > M_PI, lt = az < -M_PI;
    if (gt) az -= two_pi;
    if (lt) az += two_pi;
    double colf = ((double)W - 1.0 + 0.5) - (az + M_PI) / two_pi * (double)W;
    double cr = rint(colf);
    float rng = (float)sqrt(fma(p[2], p[2], fma(p[1], p[1], p[0] * p[0])));
    if (rows) rows[i] = row;
    if (cols) cols[i] = (int32_t)cr;
    if (ranges) ranges[i] = rng;
    if (!(cr >= 0.0 && cr < (double)W)) { bad++; continue; }
    float *px = ri + (long)row * W + (long)cr;
    if (rng < *px) *px = rng;
  }
  for (long k = 0; k < (long)H * W; k++) if (isinf(ri[k])) ri[k] = 0.0f;
  return bad;
}
