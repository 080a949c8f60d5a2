// occ_ops.cu -- A10: mmdet3d/ops/occ/occ_ops.py quantize_points / generate_dense_voxel_centers.
// The reference evaluates these with ~8 elementwise torch kernels (and a python loop over ROIs for
// the dense grids); here each is one fused kernel with the same f32 operation order.
#include "common.cuh"

namespace occb200 {

struct Wlh {
  float scale[3];
  float offset[3];
};

__global__ void k_quantize_points(const float *__restrict__ points, int64_t N, const float *__restrict__ rois,
                                  int64_t R, int roi_dim, const int64_t *__restrict__ roi_idx, float vs, Wlh w,
                                  int to_center, int64_t *__restrict__ out_coor, float *__restrict__ out_center,
                                  unsigned long long *__restrict__ n_bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * 3) return;
  const int64_t i = e / 3;
  const int j = (int)(e % 3);
  // roi_min_bound[rois_points_idx] (occ_ops.py:81) is PyTorch indexing: a negative index counts from the end,
  // anything outside [-R, R) raises.  Such a row is counted in *n_bad and gets a sentinel (never an
  // out-of-bounds read).
  int64_t ri = roi_idx[i];
  if (ri < 0) ri += R;
  if (ri < 0 || ri >= R) {
    if (j == 0 && n_bad) atomicAdd(n_bad, 1ull);
    if (to_center) out_center[e] = __int_as_float(0x7fc00000);
    else out_coor[e] = INT64_MIN;
    return;
  }
  const float *roi = rois + ri * roi_dim;
  const float size = __fadd_rn(__fmul_rn(roi[4 + j], w.scale[j]), w.offset[j]);   // occ_ops.py:76-78
  const float mn = __fdiv_rn(-size, 2.0f);                                          // :80
  const float q = floorf(__fdiv_rn(__fsub_rn(points[e], mn), vs));                  // :82
  const long long qi = (long long)q;
  if (to_center)
    out_center[e] = __fadd_rn(__fadd_rn(__fmul_rn((float)qi, vs), mn), __fdiv_rn(vs, 2.0f));   // :86-90
  else
    out_coor[e] = qi;
}

__global__ void k_dense_centers(const float *__restrict__ sizes, const int32_t *__restrict__ dims,
                                const int64_t *__restrict__ center_off, int R, int64_t total, float vs, Wlh w,
                                float *__restrict__ centers) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int lo = 0, hi = R;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (center_off[mid] <= e) lo = mid; else hi = mid;
  }
  const int r = lo;
  const int64_t f = e - center_off[r];
  const int Y = dims[3 * r + 1], Z = dims[3 * r + 2];
  const int q[3] = {(int)(f / ((int64_t)Y * Z)), (int)((f / Z) % Y), (int)(f % Z)};   // ij meshgrid (:26-36)
  for (int j = 0; j < 3; ++j) {
    const float size = __fadd_rn(__fmul_rn(sizes[3 * r + j], w.scale[j]), w.offset[j]);   // :21
    const float mn = __fdiv_rn(-size, 2.0f);                                               // :39
    centers[e * 3 + j] = __fadd_rn(__fadd_rn(__fmul_rn((float)q[j], vs), mn), __fdiv_rn(vs, 2.0f));   // :41-43
  }
}

// MirrorOccLabel (mmdet3d/datasets/pipelines/occ_pinelines.py:82-126): every UNKNOWN voxel takes the label of its
// mirror image across the grid's x mid-plane, read from the unmodified grid.  Mirror index as the reference
// computes it: long((x + 0.5 - X//2) * -1.0 + X//2), truncation toward zero (so x = X-1 maps to 0 for odd X).
__global__ void k_mirror_occ_label(const int32_t *__restrict__ labels, const int64_t *__restrict__ label_off,
                                   const int32_t *__restrict__ dims, const int32_t *__restrict__ status, int T,
                                   int32_t *__restrict__ out) {
  const int t = blockIdx.y;
  if (t >= T || (status && status[t] != 0)) return;
  const int X = dims[3 * t], Y = dims[3 * t + 1], Z = dims[3 * t + 2];
  const int64_t V = (int64_t)X * Y * Z, yz = (int64_t)Y * Z;
  const int32_t *src = labels + label_off[t];
  int32_t *dst = out + label_off[t];
  const float mid = (float)(X / 2);
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < V; f += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = src[f];
    if (v == 0) {
      const int64_t x = f / yz;
      long long mx = (long long)__fadd_rn(__fmul_rn(__fsub_rn(__fadd_rn((float)x, 0.5f), mid), -1.0f), mid);
      if (mx < 0) mx += X;                                      // PyTorch negative-index wrap (not reached)
      v = src[mx * yz + (f - x * yz)];
    }
    dst[f] = v;
  }
}

// sample_observation (mmdet3d/models/roi_heads/bbox_heads/occ_ae_head.py:100-127): the dense "observed" grid of
// every ROI in one launch.  Point i belongs to ROI roi_idx[i] (an index outside [0, R) matches no ROI, as in the
// reference's `pts_roi_inds == i` masks); its quantised coordinate marks voxel (c0 * Y + c1) * Z + c2 of that ROI's
// grid iff 0 <= c < dims (:108-111: points on the boundary are dropped).
__global__ void k_observed_labels(const int64_t *__restrict__ coors, const int64_t *__restrict__ roi_idx, int64_t N,
                                  const int32_t *__restrict__ dims, const int64_t *__restrict__ off, int64_t R,
                                  int64_t *__restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t r = roi_idx[i];
  if (r < 0 || r >= R) return;
  const int64_t c0 = coors[3 * i], c1 = coors[3 * i + 1], c2 = coors[3 * i + 2];
  const int64_t X = dims[3 * r], Y = dims[3 * r + 1], Z = dims[3 * r + 2];
  if (c0 < 0 || c1 < 0 || c2 < 0 || c0 >= X || c1 >= Y || c2 >= Z) return;
  labels[off[r] + (c0 * Y + c1) * Z + c2] = 1;                  // every writer stores 1
}

}  // namespace occb200

using namespace occb200;

extern "C" int occb200_observed_labels(const int64_t *coors, const int64_t *roi_idx, int64_t N, const int32_t *dims,
                                       const int64_t *off, int64_t R, int64_t *labels, void *stream) {
  OCC_REQUIRE(N >= 0 && R >= 0, "bad sizes");
  if (N == 0 || R == 0) return 0;
  OCC_REQUIRE(coors && roi_idx && dims && off && labels, "NULL argument");
  k_observed_labels<<<(unsigned)ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(coors, roi_idx, N, dims, off, R, labels);
  OCC_KERNEL_OK("k_observed_labels");
  return 0;
}

extern "C" int occb200_mirror_occ_label(const int32_t *labels, const int64_t *label_off, const int32_t *dims,
                                        const int32_t *status, int32_t T, int64_t max_voxels, int32_t *out,
                                        void *stream) {
  OCC_REQUIRE(T >= 0 && max_voxels >= 0, "bad sizes");
  if (T == 0 || max_voxels == 0) return 0;
  const dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_voxels, 256), 4096), (unsigned)T);
  k_mirror_occ_label<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, label_off, dims, status, T, out);
  OCC_KERNEL_OK("k_mirror_occ_label");
  return 0;
}

extern "C" int occb200_quantize_points(const float *points, int64_t N, const float *rois, int64_t R, int roi_dim,
                                       const int64_t *roi_idx, float voxel_size, const float *scale_wlh,
                                       const float *offset_wlh, int to_center, int64_t *out_coor, float *out_center,
                                       unsigned long long *n_bad, void *stream) {
  OCC_REQUIRE(N >= 0 && R >= 0 && roi_dim >= 7, "bad sizes");
  OCC_REQUIRE(to_center ? out_center != nullptr : out_coor != nullptr, "output pointer is NULL");
  if (N == 0) return 0;
  Wlh w;
  for (int j = 0; j < 3; ++j) { w.scale[j] = scale_wlh[j]; w.offset[j] = offset_wlh[j]; }
  k_quantize_points<<<(unsigned)ceil_div(N * 3, 256), 256, 0, (cudaStream_t)stream>>>(
      points, N, rois, R, roi_dim, roi_idx, voxel_size, w, to_center, out_coor, out_center, n_bad);
  OCC_KERNEL_OK("k_quantize_points");
  return 0;
}

extern "C" int occb200_dense_voxel_centers(const float *sizes, const int32_t *dims, const int64_t *center_off, int R,
                                           int64_t total, float voxel_size, const float *scale_wlh,
                                           const float *offset_wlh, float *centers, void *stream) {
  OCC_REQUIRE(R >= 0 && total >= 0, "bad sizes");
  if (R == 0 || total == 0) return 0;
  Wlh w;
  for (int j = 0; j < 3; ++j) { w.scale[j] = scale_wlh[j]; w.offset[j] = offset_wlh[j]; }
  k_dense_centers<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(sizes, dims, center_off, R, total,
                                                                                    voxel_size, w, centers);
  OCC_KERNEL_OK("k_dense_centers");
  return 0;
}
