// candidates.cu -- candidate selection from whole-frame clouds.
//
// The reference streams the full ~180 k-point cloud of a frame through points_in_boxes_gpu once per
// (tracklet, frame) and compacts it with a boolean mask (tools/occ/occ_annotate.py:96-112; the cloud is the
// KITTI-format .bin of tools/ctrl/utils.py:60-66).  Here every frame cloud is read ONCE: a CTA takes 2 048
// consecutive points of one frame and tests each against the candidate spheres of all the boxes alive in that
// frame (shared memory); the points inside a sphere -- a superset of the box's in-box points, the exact test is
// k_crop_voxelize's -- are appended to that tracklet-frame's candidate list in cloud order.
//   pass 1 (out_points == NULL)  counts[f][k][chunk][warp] = hits of warp `warp` of chunk `chunk` in sphere k of frame f
//   (the caller turns the counts into an exclusive prefix sum, same layout)
//   pass 2                       the same tests; each warp writes its hits at its prefix, lanes in point order
// The layout [frame][box][chunk][warp] makes every tracklet-frame's candidates contiguous and in cloud order.
#include "common.cuh"

namespace occb200 {

constexpr int kCandThreads = 256;
constexpr int kCandPerWarp = 256;                       // consecutive points per warp
constexpr int kCandChunk = kCandPerWarp * (kCandThreads / 32);
constexpr int kCandBoxTile = 256;

__global__ void __launch_bounds__(kCandThreads)
k_select_candidates(const float *__restrict__ clouds, int stride, const int64_t *__restrict__ cloud_off,
                    const occb200_cand_box_t *__restrict__ boxes, const int64_t *__restrict__ frame_box_off,
                    const int64_t *__restrict__ cnt_off, int32_t *__restrict__ counts,
                    const int64_t *__restrict__ prefix, float *__restrict__ out_points, int out_stride) {
  __shared__ float4 s_box[kCandBoxTile];
  const int f = blockIdx.y;
  const int64_t p0 = cloud_off[f], p1 = cloud_off[f + 1];
  const int64_t nchunk = (p1 - p0 + kCandChunk - 1) / kCandChunk;
  const int chunk = blockIdx.x;
  if (chunk >= nchunk) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b0 = frame_box_off[f], K = frame_box_off[f + 1] - b0;
  const int64_t wbase = p0 + (int64_t)chunk * kCandChunk + (int64_t)warp * kCandPerWarp;
  for (int64_t t0 = 0; t0 < K; t0 += kCandBoxTile) {
    const int nt = (int)min((int64_t)kCandBoxTile, K - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < nt; k += kCandThreads) {
      const occb200_cand_box_t b = boxes[b0 + t0 + k];
      s_box[k] = make_float4(b.cx, b.cy, b.cz, b.r2);
    }
    __syncthreads();
    // running write position of this warp per box of the tile: held by lane (k & 31) of round (k >> 5)
    for (int kr = 0; kr < nt; kr += 32) {
      const int kmine = kr + lane;
      int64_t pos = 0;
      int cnt = 0;
      if (out_points && kmine < nt)
        pos = prefix[cnt_off[f] + (((t0 + kmine) * nchunk + chunk) * (kCandThreads / 32) + warp)];
      for (int i = 0; i < kCandPerWarp; i += 32) {
        const int64_t j = wbase + i + lane;
        const bool ok = j < p1;
        float x = 0.f, y = 0.f, z = 0.f;
        if (ok) {
          const float *p = clouds + j * stride;
          x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
        }
        const int kend = min(32, nt - kr);
        for (int kk = 0; kk < kend; ++kk) {
          const float4 b = s_box[kr + kk];
          const float dx = x - b.x, dy = y - b.y, dz = z - b.z;
          const bool hit = ok && (dx * dx + dy * dy + dz * dz <= b.w);
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (m == 0u) continue;
          if (out_points) {
            const int64_t base = __shfl_sync(0xffffffffu, pos, kk) + __shfl_sync(0xffffffffu, cnt, kk);
            if (hit) {
              float *o = out_points + (base + __popc(m & ((1u << lane) - 1u))) * out_stride;
              const float *p = clouds + j * stride;
              for (int c = 0; c < out_stride; ++c) o[c] = __ldg(p + c);
            }
          }
          if (lane == kk) cnt += __popc(m);
        }
      }
      if (!out_points && kmine < nt)
        counts[cnt_off[f] + (((t0 + kmine) * nchunk + chunk) * (kCandThreads / 32) + warp)] = cnt;
    }
  }
}

}  // namespace occb200

using namespace occb200;

extern "C" int occb200_candidate_chunk(void) { return kCandChunk; }
extern "C" int occb200_candidate_warps(void) { return kCandThreads / 32; }

extern "C" int occb200_select_candidates(const float *clouds, int stride, const int64_t *cloud_off, int32_t NF,
                                         int64_t max_cloud, const occb200_cand_box_t *boxes,
                                         const int64_t *frame_box_off, const int64_t *cnt_off, int32_t *counts,
                                         const int64_t *prefix, float *out_points, int out_stride, void *stream) {
  OCC_REQUIRE(NF >= 0 && stride >= 3 && max_cloud >= 0, "bad sizes");
  OCC_REQUIRE(out_points == nullptr || (prefix != nullptr && out_stride >= 3 && out_stride <= stride),
              "pass 2 needs the prefix sums and 3 <= out_stride <= stride");
  OCC_REQUIRE(out_points != nullptr || counts != nullptr, "pass 1 needs counts");
  if (NF == 0 || max_cloud == 0) return 0;
  const dim3 grid((unsigned)ceil_div(max_cloud, kCandChunk), (unsigned)NF);
  k_select_candidates<<<grid, kCandThreads, 0, (cudaStream_t)stream>>>(clouds, stride, cloud_off, boxes, frame_box_off,
                                                                       cnt_off, counts, prefix, out_points, out_stride);
  OCC_KERNEL_OK("k_select_candidates");
  return 0;
}
