// Range-image builder: the `*_RANGE_IMAGE_MERGE_VIRTUAL` step of the reference's converter
// (tools/data_converter/waymo_converter.py:632-670), which calls
// waymo_open_dataset.utils.range_image_utils.build_range_image_from_point_cloud (waymo-open-dataset-tf-2-1-0
// == 1.2.0, requirements/optional.txt; NOT vendored under the reference -- its published algorithm is restated
// here and in oracle/oracle.py::build_range_image; parity unpinned against the TF implementation itself).
//
//   p      = R x + t            (R, t from inv(extrinsic) in f64, host)        -- einsum 'bij,bkj->bik' + translation
//   row    = argmin_h |atan2(p.z, |p.xy|) - incl[h]|                             (first index on ties)
//   az     = atan2(p.y, p.x) + atan2(E[1,0], E[0,0]);  +-2 pi if outside [-pi, pi]
//   col    = int32(round_half_even(W - 1 + 0.5 - (az + pi) / (2 pi) * W))        (asserted in [0, W) by TF)
//   ri[row, col] = min over points of f32(|p|);  0 where no point lands
// It is the forward use of the projection the ray-cast inverts (occ_annotate.py:141-201 is a torch port of the
// first half of the same function), so the exact f64 helpers of geom.cuh are reused.
#include "common.cuh"
#include "geom.cuh"

namespace occb200 {

__global__ void __launch_bounds__(256)
k_ri_scatter(const float *__restrict__ points, int stride, const int64_t *__restrict__ pt_off,
             const occb200_ri_desc_t *__restrict__ desc, const float *__restrict__ incl_pool,
             unsigned int *__restrict__ img_pool, unsigned long long *__restrict__ n_bad) {
  const int b = blockIdx.y;
  const occb200_ri_desc_t d = desc[b];
  const float *incl = incl_pool + d.incl_off;
  unsigned int *img = img_pool + d.ri_off;
  const double kPi = 3.14159265358979323846, kTwoPi = 6.28318530717958647692;
  unsigned long long bad = 0;
  for (int64_t j = pt_off[b] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < pt_off[b + 1];
       j += (int64_t)gridDim.x * blockDim.x) {
    const double x = (double)points[j * stride], y = (double)points[j * stride + 1], z = (double)points[j * stride + 2];
    const double px = __dadd_rn(__fma_rn(z, d.v2l[2], __fma_rn(y, d.v2l[1], __dmul_rn(x, d.v2l[0]))), d.v2l[3]);
    const double py = __dadd_rn(__fma_rn(z, d.v2l[6], __fma_rn(y, d.v2l[5], __dmul_rn(x, d.v2l[4]))), d.v2l[7]);
    const double pz = __dadd_rn(__fma_rn(z, d.v2l[10], __fma_rn(y, d.v2l[9], __dmul_rn(x, d.v2l[8]))), d.v2l[11]);
    const double xy = __dsqrt_rn(__fma_rn(py, py, __dmul_rn(px, px)));
    const int row = nearest_row(atan2(pz, xy), incl, d.H, d.mono);
    double az = __dadd_rn(atan2(py, px), d.azc);
    const bool gt = az > kPi, lt = az < -kPi;
    if (gt) az = __dsub_rn(az, kTwoPi);
    if (lt) az = __dadd_rn(az, kTwoPi);
    const double w = (double)d.W;
    const double colf = __dsub_rn(__dadd_rn(__dsub_rn(w, 1.0), 0.5), __dmul_rn(__ddiv_rn(__dadd_rn(az, kPi), kTwoPi), w));
    const double cr = rint(colf);
    if (!(cr >= 0.0 && cr < w)) {                 // TF: assert_non_negative / assert_less
      ++bad;
      continue;
    }
    const float rng = (float)__dsqrt_rn(__fma_rn(pz, pz, __fma_rn(py, py, __dmul_rn(px, px))));
    // ranges are >= 0: their bit patterns order like the values
    atomicMin(img + (int64_t)row * d.W + (int)cr, __float_as_uint(rng));
  }
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, bad);
}

__global__ void k_ri_finalize(unsigned int *__restrict__ img_pool, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (img_pool[i] == 0xffffffffu) img_pool[i] = 0u;      // no point landed here: 0.0f
}

}  // namespace occb200

using namespace occb200;

extern "C" int occb200_build_range_images(const float *points, int point_stride, const int64_t *pt_off,
                                          int64_t max_points, const occb200_ri_desc_t *desc, int32_t n_images,
                                          const float *incl_pool, float *ri_pool, int64_t ri_len,
                                          unsigned long long *n_bad, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(n_images >= 0 && ri_len >= 0 && point_stride >= 3 && max_points >= 0, "bad sizes");
  OCC_REQUIRE(n_bad != nullptr, "n_bad is NULL");
  OCC_CUDA(cudaMemsetAsync(n_bad, 0, 8, stream));
  if (n_images == 0 || ri_len == 0) return 0;
  OCC_CUDA(cudaMemsetAsync(ri_pool, 0xff, 4 * (size_t)ri_len, stream));
  if (max_points > 0) {
    const dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_points, 256), kNumSMs * 8), (unsigned)n_images);
    k_ri_scatter<<<grid, 256, 0, stream>>>(points, point_stride, pt_off, desc, incl_pool, (unsigned int *)ri_pool, n_bad);
    OCC_KERNEL_OK("k_ri_scatter");
  }
  k_ri_finalize<<<(unsigned)std::min<int64_t>(ceil_div(ri_len, 256), kNumSMs * 8), 256, 0, stream>>>(
      (unsigned int *)ri_pool, ri_len);
  OCC_KERNEL_OK("k_ri_finalize");
  return 0;
}
