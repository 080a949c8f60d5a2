// points_in_boxes.cu -- A1: which box is each point in.
//
// Replaces mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:51-105 (one thread per point,
// boxes re-read from global for every point, 3 strided scalar loads).  Here: the batch's boxes are
// turned into BoxTest records once per CTA in shared memory (trig evaluated once per box instead of
// once per point-box pair), each thread owns 4 consecutive points (three 16-byte loads when the row
// is aligned), and the index row is written with one 16-byte store.
#include "common.cuh"
#include "geom.cuh"

namespace occb200 {

constexpr int kPibThreads = 256;
constexpr int kPibPerThread = 4;
constexpr int kPibBoxTile = 256;   // boxes staged per pass

template <bool kBatch>
__global__ void __launch_bounds__(kPibThreads)
k_points_in_boxes(const float *__restrict__ boxes, const float *__restrict__ pts, const float *__restrict__ trig,
                  int32_t *__restrict__ out, int T, int M, int contract) {
  __shared__ BoxTest s_box[kPibBoxTile];
  const int b = blockIdx.y;
  const float *bx = boxes + (int64_t)b * T * 7;
  const float *pp = pts + (int64_t)b * M * 3;
  const int64_t m0 = ((int64_t)blockIdx.x * kPibThreads + threadIdx.x) * kPibPerThread;
  float x[kPibPerThread], y[kPibPerThread], z[kPibPerThread];
  int hit[kPibPerThread];
  const bool full = m0 + kPibPerThread <= M;
  if (full && ((reinterpret_cast<uintptr_t>(pp + m0 * 3) & 15) == 0)) {
    const float4 *v = reinterpret_cast<const float4 *>(pp + m0 * 3);
    const float4 a = __ldg(v), c = __ldg(v + 1), d = __ldg(v + 2);
    x[0] = a.x; y[0] = a.y; z[0] = a.z;
    x[1] = a.w; y[1] = c.x; z[1] = c.y;
    x[2] = c.z; y[2] = c.w; z[2] = d.x;
    x[3] = d.y; y[3] = d.z; z[3] = d.w;
  } else {
#pragma unroll
    for (int k = 0; k < kPibPerThread; ++k) {
      const bool ok = m0 + k < M;
      x[k] = ok ? __ldg(pp + (m0 + k) * 3) : 0.f;
      y[k] = ok ? __ldg(pp + (m0 + k) * 3 + 1) : 0.f;
      z[k] = ok ? __ldg(pp + (m0 + k) * 3 + 2) : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < kPibPerThread; ++k) hit[k] = -1;

  for (int t0 = 0; t0 < T; t0 += kPibBoxTile) {
    const int nt = min(kPibBoxTile, T - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < nt; k += kPibThreads) {
      const float *bk = bx + (int64_t)(t0 + k) * 7;
      float ca, sa;
      if (trig) {
        ca = trig[((int64_t)b * T + t0 + k) * 2];
        sa = trig[((int64_t)b * T + t0 + k) * 2 + 1];
      } else {
        const float a = box_rot_angle(bk[6]);   // same libdevice cosf/sinf as the reference kernel
        ca = cosf(a);
        sa = sinf(a);
      }
      s_box[k] = make_box_test(bk, ca, sa, contract);
    }
    __syncthreads();
    if (kBatch) {
      for (int k = 0; k < nt; ++k) {
        const BoxTest bt = s_box[k];
#pragma unroll
        for (int q = 0; q < kPibPerThread; ++q)
          if (m0 + q < M) out[((int64_t)b * M + m0 + q) * T + t0 + k] = pt_in_box(bt, x[q], y[q], z[q]) ? 1 : 0;
      }
    } else {
      for (int k = 0; k < nt; ++k) {
        const BoxTest bt = s_box[k];
#pragma unroll
        for (int q = 0; q < kPibPerThread; ++q)
          if (hit[q] < 0 && pt_in_box(bt, x[q], y[q], z[q])) hit[q] = t0 + k;   // first hit wins (:69-75)
      }
    }
  }
  if (!kBatch) {
    int32_t *o = out + (int64_t)b * M + m0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      *reinterpret_cast<int4 *>(o) = make_int4(hit[0], hit[1], hit[2], hit[3]);
    } else {
#pragma unroll
      for (int q = 0; q < kPibPerThread; ++q)
        if (m0 + q < M) o[q] = hit[q];
    }
  }
}

template <bool kBatch>
static int launch_pib(const float *boxes, const float *pts, const float *trig, int32_t *out, int B, int T, int M,
                      int contract, cudaStream_t stream) {
  OCC_REQUIRE(B >= 0 && T >= 0 && M >= 0, "negative size");
  if (B == 0 || M == 0) return 0;
  dim3 grid((unsigned)ceil_div(M, kPibThreads * kPibPerThread), (unsigned)B);
  k_points_in_boxes<kBatch><<<grid, kPibThreads, 0, stream>>>(boxes, pts, trig, out, T, M, contract);
  OCC_KERNEL_OK("k_points_in_boxes");
  return 0;
}

}  // namespace occb200

extern "C" int occb200_points_in_boxes_gpu(const float *boxes, const float *pts, const float *trig, int32_t *out,
                                           int B, int T, int M, int contract, void *stream) {
  return occb200::launch_pib<false>(boxes, pts, trig, out, B, T, M, contract, (cudaStream_t)stream);
}

extern "C" int occb200_points_in_boxes_batch(const float *boxes, const float *pts, const float *trig, int32_t *out,
                                             int B, int T, int M, int contract, void *stream) {
  return occb200::launch_pib<true>(boxes, pts, trig, out, B, T, M, contract, (cudaStream_t)stream);
}
