// point_pool.cu -- dynamic_point_pool_mixed: the points of every ROI (box + margin) with ROI-local coordinates.
//
// Reference call site: TrackletPointRoIExtractor (mmdet3d/models/roi_heads/roi_extractors/
// dynamic_point_roi_extractor.py:176-300) -> mmdet3d/ops/dynamic_point_pool_op.py:63-113 ->
// dynamic_point_pool_ext.dynamic_point_pool_mixed_gpu.  The extension's source is NOT in the reference tree (it comes
// from TorchEx / the SST code base, un-vendored, no pinned version): PARITY UNPINNED.  What is restated here is
// what the reference itself checks about the op's output (the debug assertions of the extractor, :214-230) plus the
// box convention of this code base's own in-box kernel (points_in_boxes_cuda.cu:24-49):
//   a = rz + pi/2;  local_x = dx cos a - dy sin a  (extent roi[4]),  local_y = dx sin a + dy cos a  (extent roi[3]),
//   local_z = z - (z_bottom + h/2)  (extent roi[5]);   a point belongs to the ROI iff it lies strictly inside the box
//   enlarged by extra_wlh (extra[0] on local x, extra[1] on local y, extra[2] on z) and has the ROI's batch index;
//   feats[13] = xyz, local xyz, (local + size/2), (size/2 - local), is_in_margin (inside the enlarged box only).
// The reference kernel appends hits in atomic order ("not strictly guaranteed" sorted, dynamic_point_pool_op.py:36)
// and truncates at max_inbox_point per ROI / max_all_pts overall in that order; here the order is (ROI, point index)
// and the truncations keep the lowest point indices: deterministic.
//
// Two passes over (ROI, points of its batch index) -- count, then write at per-ROI cursors -- the caller sorts the
// hits of a ROI by point index and applies the truncations (point_pool.py).
#include "common.cuh"

namespace occb200 {

struct PoolBox {
  float cx, cy, cz, hl, hw, hh, el, ew, eh, cosa, sina;   // centre, half sizes, enlarged half sizes, rotation
};

__device__ __forceinline__ PoolBox pool_box(const float *__restrict__ r, float e0, float e1, float e2) {
  PoolBox b;
  b.cx = r[0]; b.cy = r[1];
  b.hw = 0.5f * r[3]; b.hl = 0.5f * r[4]; b.hh = 0.5f * r[5];
  b.cz = r[2] + b.hh;
  b.el = b.hl + 0.5f * e0; b.ew = b.hw + 0.5f * e1; b.eh = b.hh + 0.5f * e2;
  const float a = r[6] + 1.5707963267948966f;
  b.cosa = cosf(a); b.sina = sinf(a);
  return b;
}

template <bool WRITE>
__global__ void __launch_bounds__(256)
k_point_pool(const float *__restrict__ rois, const int64_t *__restrict__ roi_lo, const int64_t *__restrict__ roi_hi,
             const float *__restrict__ pts, const int64_t *__restrict__ perm, float e0, float e1, float e2,
             int32_t *__restrict__ counts, const int64_t *__restrict__ base, int32_t *__restrict__ cursor,
             int64_t *__restrict__ out_pidx, int64_t *__restrict__ out_roi, float *__restrict__ out_feats) {
  const int r = blockIdx.y;
  const int64_t lo = roi_lo[r], hi = roi_hi[r];
  const PoolBox b = pool_box(rois + 7 * (int64_t)r, e0, e1, e2);
  int local = 0;
  for (int64_t j = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < hi; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = perm[j];
    const float x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
    const float dx = x - b.cx, dy = y - b.cy, lz = z - b.cz;
    const float lx = dx * b.cosa - dy * b.sina, ly = dx * b.sina + dy * b.cosa;
    const bool in_large = fabsf(lz) < b.eh && fabsf(lx) < b.el && fabsf(ly) < b.ew;
    if (!in_large) continue;
    if (!WRITE) { ++local; continue; }
    const bool in_small = fabsf(lz) < b.hh && fabsf(lx) < b.hl && fabsf(ly) < b.hw;
    const int64_t o = base[r] + atomicAdd(cursor + r, 1);
    out_pidx[o] = p;
    out_roi[o] = r;
    float *f = out_feats + 13 * o;
    f[0] = x; f[1] = y; f[2] = z; f[3] = lx; f[4] = ly; f[5] = lz;
    f[6] = lx + b.hl; f[7] = ly + b.hw; f[8] = lz + b.hh;
    f[9] = b.hl - lx; f[10] = b.hw - ly; f[11] = b.hh - lz;
    f[12] = in_small ? 0.f : 1.f;
  }
  if (!WRITE) {
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(counts + r, local);
  }
}

}  // namespace occb200

using namespace occb200;

extern "C" int occb200_point_pool(const float *rois, const int64_t *roi_lo, const int64_t *roi_hi, int64_t R,
                                  const float *pts, const int64_t *perm, int64_t max_range, const float *extra_wlh,
                                  int32_t *counts, const int64_t *base, int32_t *cursor, int64_t *out_pidx,
                                  int64_t *out_roi, float *out_feats, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(R >= 0 && max_range >= 0, "bad sizes");
  if (R == 0 || max_range == 0) return 0;
  OCC_REQUIRE(R <= 65535, "more than 65535 ROIs per call");
  const dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_range, 256), 64), (unsigned)R);
  if (out_feats == nullptr) {
    OCC_REQUIRE(counts != nullptr, "pass 1 needs counts");
    k_point_pool<false><<<grid, 256, 0, stream>>>(rois, roi_lo, roi_hi, pts, perm, extra_wlh[0], extra_wlh[1], extra_wlh[2],
                                                  counts, nullptr, nullptr, nullptr, nullptr, nullptr);
  } else {
    OCC_REQUIRE(base && cursor && out_pidx && out_roi, "pass 2 needs base, cursor and the outputs");
    k_point_pool<true><<<grid, 256, 0, stream>>>(rois, roi_lo, roi_hi, pts, perm, extra_wlh[0], extra_wlh[1], extra_wlh[2],
                                                 nullptr, base, cursor, out_pidx, out_roi, out_feats);
  }
  OCC_KERNEL_OK("k_point_pool");
  return 0;
}
