// voxelize.cu -- A7 dynamic voxelization and A8 hard voxelization.
//
// Replaces mmdet3d/ops/voxel/src/voxelization_cuda.cu:
//   dynamic (:24-65, :332-375): 64-thread blocks, three strided scalar loads and three strided
//     stores per point, cudaDeviceSynchronize per call.  Here: a CTA stages 256 whole rows with
//     16-byte coalesced loads, and writes its 768 coordinates with 16-byte stores; no sync.
//   hard (:67-184, :188-330): O(N^2) duplicate scan + a <<<1,1>>> serial pass + 4 device syncs.
//     Here: stable radix sort by cell id (CUB), segment heads -> first-appearance order by a
//     prefix sum over original indices, one gather kernel; one sync (the count is returned).
// Semantics follow the CPU implementation voxelization_cpu.cpp (clamping fork, zyx order).
#include <cub/cub.cuh>

#include "common.cuh"

namespace occb200 {

constexpr int kVoxThreads = 256;
constexpr int kVoxMaxStageCols = 16;

struct VoxParams {
  float vs[3];
  float mn[3];
  int grid[3];
};

template <typename T>
__device__ __forceinline__ int quantise(T p, float mn, float vs, int grid);

template <>
__device__ __forceinline__ int quantise<float>(float p, float mn, float vs, int grid) {
  // c = floor((p - min) / vs) in f32 (voxelization_cpu.cpp:22), clamped (:25-30)
  int c = (int)floorf(__fdiv_rn(__fsub_rn(p, mn), vs));
  return c < 0 ? 0 : (c >= grid ? grid - 1 : c);
}
template <>
__device__ __forceinline__ int quantise<double>(double p, float mn, float vs, int grid) {
  int c = (int)floor(__ddiv_rn(__dsub_rn(p, (double)mn), (double)vs));
  return c < 0 ? 0 : (c >= grid ? grid - 1 : c);
}

// f32 rows staged through shared memory (C <= 16, 16-byte aligned base)
__global__ void __launch_bounds__(kVoxThreads)
k_dynamic_voxelize_staged(const float *__restrict__ points, int64_t N, int C, VoxParams pr,
                          int32_t *__restrict__ coors) {
  __shared__ __align__(16) float s_rows[kVoxThreads * kVoxMaxStageCols];
  __shared__ __align__(16) int32_t s_out[kVoxThreads * 3];
  const int64_t row0 = (int64_t)blockIdx.x * kVoxThreads;
  const int rows = (int)min((int64_t)kVoxThreads, N - row0);
  const int nfl = rows * C;
  const float *src = points + row0 * C;
  const int nvec = nfl >> 2;
  for (int v = threadIdx.x; v < nvec; v += kVoxThreads)
    reinterpret_cast<float4 *>(s_rows)[v] = __ldg(reinterpret_cast<const float4 *>(src) + v);
  for (int e = (nvec << 2) + threadIdx.x; e < nfl; e += kVoxThreads) s_rows[e] = __ldg(src + e);
  __syncthreads();
  if (threadIdx.x < rows) {
    const float *p = s_rows + threadIdx.x * C;
    s_out[threadIdx.x * 3 + 0] = quantise<float>(p[2], pr.mn[2], pr.vs[2], pr.grid[2]);
    s_out[threadIdx.x * 3 + 1] = quantise<float>(p[1], pr.mn[1], pr.vs[1], pr.grid[1]);
    s_out[threadIdx.x * 3 + 2] = quantise<float>(p[0], pr.mn[0], pr.vs[0], pr.grid[0]);
  }
  __syncthreads();
  int32_t *dst = coors + row0 * 3;
  const int nout = rows * 3;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int ov = nout >> 2;
    for (int v = threadIdx.x; v < ov; v += kVoxThreads)
      reinterpret_cast<int4 *>(dst)[v] = reinterpret_cast<const int4 *>(s_out)[v];
    for (int e = (ov << 2) + threadIdx.x; e < nout; e += kVoxThreads) dst[e] = s_out[e];
  } else {
    for (int e = threadIdx.x; e < nout; e += kVoxThreads) dst[e] = s_out[e];
  }
}

template <typename T>
__global__ void __launch_bounds__(kVoxThreads)
k_dynamic_voxelize_direct(const T *__restrict__ points, int64_t N, int C, VoxParams pr,
                          int32_t *__restrict__ coors) {
  const int64_t i = (int64_t)blockIdx.x * kVoxThreads + threadIdx.x;
  if (i >= N) return;
  const T *p = points + i * C;
  coors[i * 3 + 0] = quantise<T>(p[2], pr.mn[2], pr.vs[2], pr.grid[2]);
  coors[i * 3 + 1] = quantise<T>(p[1], pr.mn[1], pr.vs[1], pr.grid[1]);
  coors[i * 3 + 2] = quantise<T>(p[0], pr.mn[0], pr.vs[0], pr.grid[0]);
}

static VoxParams make_params(const float *voxel_size, const float *coors_range, bool round_grid) {
  VoxParams pr;
  for (int i = 0; i < 3; ++i) {
    pr.vs[i] = voxel_size[i];
    pr.mn[i] = coors_range[i];
    const float g = (coors_range[3 + i] - coors_range[i]) / voxel_size[i];
    pr.grid[i] = round_grid ? (int)roundf(g) : (int)ceilf(g);   // voxelization_cpu.cpp:125 / :158
  }
  return pr;
}

// ----------------------------------------------------------------------------- hard voxelize
__global__ void k_hv_keys(const float *__restrict__ points, int64_t N, int C, VoxParams pr,
                          uint64_t *__restrict__ keys, int32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float *p = points + i * C;
  const int cx = quantise<float>(p[0], pr.mn[0], pr.vs[0], pr.grid[0]);
  const int cy = quantise<float>(p[1], pr.mn[1], pr.vs[1], pr.grid[1]);
  const int cz = quantise<float>(p[2], pr.mn[2], pr.vs[2], pr.grid[2]);
  keys[i] = ((uint64_t)cz * pr.grid[1] + cy) * pr.grid[0] + cx;
  idx[i] = (int32_t)i;
}

__global__ void k_hv_heads(const uint64_t *__restrict__ keys, const int32_t *__restrict__ idx, int64_t N,
                           int32_t *__restrict__ is_first, int32_t *__restrict__ seg_start) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const bool head = (j == 0) || keys[j] != keys[j - 1];
  seg_start[j] = head ? (int32_t)j : 0;
  if (head) is_first[idx[j]] = 1;      // stable sort: the head is the voxel's first point
}

__global__ void k_hv_gather(const float *__restrict__ points, int64_t N, int C, VoxParams pr,
                            const uint64_t *__restrict__ keys, const int32_t *__restrict__ idx,
                            const int32_t *__restrict__ seg_start, const int32_t *__restrict__ first_rank,
                            int max_points, int max_voxels, float *__restrict__ voxels,
                            int32_t *__restrict__ coors, int32_t *__restrict__ num_points) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int s = seg_start[j];
  const int rank = (int)(j - s);
  const int vid = first_rank[idx[s]];            // voxels numbered in first-appearance order
  if (max_voxels != -1 && vid >= max_voxels) return;
  if (rank == 0) {
    const uint64_t k = keys[j];
    coors[vid * 3 + 2] = (int32_t)(k % pr.grid[0]);
    coors[vid * 3 + 1] = (int32_t)((k / pr.grid[0]) % pr.grid[1]);
    coors[vid * 3 + 0] = (int32_t)(k / ((uint64_t)pr.grid[0] * pr.grid[1]));
  }
  if (max_points != -1 && rank >= max_points) return;
  const float *p = points + (int64_t)idx[j] * C;
  float *dst = voxels + ((int64_t)vid * max_points + rank) * C;
  for (int k = 0; k < C; ++k) dst[k] = p[k];
  atomicAdd(&num_points[vid], 1);
}

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

struct HvLayout {
  int64_t keys_a, keys_b, idx_a, idx_b, is_first, first_rank, seg_start, total, cub, cub_bytes, bytes;
};

static HvLayout hv_layout(int64_t N) {
  HvLayout l;
  int64_t off = 0;
  auto take = [&](int64_t b) { int64_t o = off; off = align_up(off + b, 256); return o; };
  const int64_t n = N > 0 ? N : 1;
  l.keys_a = take(8 * n); l.keys_b = take(8 * n);
  l.idx_a = take(4 * n); l.idx_b = take(4 * n);
  l.is_first = take(4 * n); l.first_rank = take(4 * n); l.seg_start = take(4 * n);
  l.total = take(8);
  size_t s1 = 0, s2 = 0, s3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, s1, (uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, s2, (int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  cub::DeviceScan::InclusiveScan(nullptr, s3, (int32_t *)nullptr, (int32_t *)nullptr, MaxOp(), (int)n);
  l.cub_bytes = (int64_t)std::max(s1, std::max(s2, s3));
  l.cub = take(l.cub_bytes);
  l.bytes = off;
  return l;
}

}  // namespace occb200

using namespace occb200;

extern "C" int occb200_dynamic_voxelize(const void *points, int dtype, int64_t N, int C, const float *voxel_size,
                                        const float *coors_range, int32_t *coors, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(N >= 0 && C >= 3, "need N >= 0 and at least 3 columns");
  OCC_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (f32) or 1 (f64)");
  if (N == 0) return 0;
  const VoxParams pr = make_params(voxel_size, coors_range, false);
  const unsigned grid = (unsigned)ceil_div(N, kVoxThreads);
  if (dtype == 0) {
    if (C <= kVoxMaxStageCols && (reinterpret_cast<uintptr_t>(points) & 15) == 0) {
      k_dynamic_voxelize_staged<<<grid, kVoxThreads, 0, stream>>>((const float *)points, N, C, pr, coors);
      OCC_KERNEL_OK("k_dynamic_voxelize_staged");
    } else {
      k_dynamic_voxelize_direct<float><<<grid, kVoxThreads, 0, stream>>>((const float *)points, N, C, pr, coors);
      OCC_KERNEL_OK("k_dynamic_voxelize_direct<float>");
    }
  } else {
    k_dynamic_voxelize_direct<double><<<grid, kVoxThreads, 0, stream>>>((const double *)points, N, C, pr, coors);
    OCC_KERNEL_OK("k_dynamic_voxelize_direct<double>");
  }
  return 0;
}

extern "C" int64_t occb200_hard_voxelize_workspace_bytes(int64_t N) { return hv_layout(N).bytes; }

extern "C" int occb200_hard_voxelize(const float *points, int64_t N, int C, const float *voxel_size,
                                     const float *coors_range, int max_points, int max_voxels, float *voxels,
                                     int32_t *coors, int32_t *num_points_per_voxel, void *workspace,
                                     int64_t workspace_bytes, int *voxel_num_host, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(N >= 0 && C >= 3, "need N >= 0 and at least 3 columns");
  OCC_REQUIRE(N < (1ll << 31), "N must fit int32");
  OCC_REQUIRE(voxel_num_host != nullptr, "voxel_num_host is NULL");
  *voxel_num_host = 0;
  if (N == 0) return 0;
  const HvLayout l = hv_layout(N);
  OCC_REQUIRE(workspace != nullptr && workspace_bytes >= l.bytes, "workspace too small");
  const VoxParams pr = make_params(voxel_size, coors_range, true);
  char *ws = (char *)workspace;
  uint64_t *keys_a = (uint64_t *)(ws + l.keys_a), *keys_b = (uint64_t *)(ws + l.keys_b);
  int32_t *idx_a = (int32_t *)(ws + l.idx_a), *idx_b = (int32_t *)(ws + l.idx_b);
  int32_t *is_first = (int32_t *)(ws + l.is_first), *first_rank = (int32_t *)(ws + l.first_rank);
  int32_t *seg_start = (int32_t *)(ws + l.seg_start);
  const unsigned grid = (unsigned)ceil_div(N, 256);
  k_hv_keys<<<grid, 256, 0, stream>>>(points, N, C, pr, keys_a, idx_a);
  OCC_KERNEL_OK("k_hv_keys");
  const double cells = (double)pr.grid[0] * pr.grid[1] * pr.grid[2];
  int end_bit = 1;
  while (end_bit < 64 && (double)(1ull << end_bit) < cells) ++end_bit;
  size_t cb = (size_t)l.cub_bytes;
  OCC_CUDA(cub::DeviceRadixSort::SortPairs(ws + l.cub, cb, keys_a, keys_b, idx_a, idx_b, (int)N, 0, end_bit, stream));
  count_launch(3);
  OCC_CUDA(cudaMemsetAsync(is_first, 0, 4 * N, stream));
  k_hv_heads<<<grid, 256, 0, stream>>>(keys_b, idx_b, N, is_first, seg_start);
  OCC_KERNEL_OK("k_hv_heads");
  cb = (size_t)l.cub_bytes;
  OCC_CUDA(cub::DeviceScan::ExclusiveSum(ws + l.cub, cb, is_first, first_rank, (int)N, stream));
  cb = (size_t)l.cub_bytes;
  OCC_CUDA(cub::DeviceScan::InclusiveScan(ws + l.cub, cb, seg_start, seg_start, MaxOp(), (int)N, stream));
  count_launch(4);
  k_hv_gather<<<grid, 256, 0, stream>>>(points, N, C, pr, keys_b, idx_b, seg_start, first_rank, max_points,
                                        max_voxels, voxels, coors, num_points_per_voxel);
  OCC_KERNEL_OK("k_hv_gather");
  int32_t tail[2] = {0, 0};
  OCC_CUDA(cudaMemcpyAsync(&tail[0], first_rank + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
  OCC_CUDA(cudaMemcpyAsync(&tail[1], is_first + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
  OCC_CUDA(cudaStreamSynchronize(stream));
  int total = tail[0] + tail[1];
  if (max_voxels != -1 && total > max_voxels) total = max_voxels;
  *voxel_num_host = total;
  return 0;
}
