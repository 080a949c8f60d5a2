// scatter.cu -- A6 / A9: sorted-unique voxel rows and deterministic segmented reductions.
//
// Replaces
//   mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-234 (at::unique_dim + one thread per point
//     doing C serial float atomics / CAS-max into [M,C]) and :236-303 (backward), and
//   mmdet3d/ops/sst/sst_ops.py:150-181 (torch.unique(dim=0) + torch_scatter).
// Design: pack each coordinate row into ONE 64-bit key (per-column bit widths from a min/max
// pass), stable radix sort of (key, point index), segment heads -> unique rows / inverse / counts
// in the reference's lexicographic order, and a reduction plan (order, gstart).  Reductions walk a
// voxel's points in ascending index with one (sub-)warp per voxel: no atomics, bit-reproducible.
#include <cub/cub.cuh>

#include "common.cuh"

namespace occb200 {

constexpr int kMaxK = 8;

struct KeySpec {
  long long bias[kMaxK];   // subtracted from each column
  int shift[kMaxK];        // left shift of each column inside the key
  int K;
  int mode;
  int flag_shift;          // modes 1/2: position of the "valid" flag bit
  long long batch_size;    // mode 2
  long long vmax[kMaxK];   // bounded call: largest value a valid row may hold in the column (LLONG_MAX: unbounded)
  unsigned long long *oob; // bounded call: set to 1 by a valid row beyond vmax (the caller then repeats the call unbounded)
};

template <typename CT>
__global__ void k_minmax(const CT *__restrict__ coors, int64_t N, int K, long long *__restrict__ mm) {
  // mm[2k] = min of column k, mm[2k+1] = max
  long long lo[kMaxK], hi[kMaxK];
  for (int k = 0; k < kMaxK; ++k) { lo[k] = LLONG_MAX; hi[k] = LLONG_MIN; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < K; ++k) {
      const long long v = (long long)coors[i * K + k];
      lo[k] = v < lo[k] ? v : lo[k];
      hi[k] = v > hi[k] ? v : hi[k];
    }
  // warp -> CTA -> ONE atomic pair per column and CTA (per-warp atomics on the 2K result words serialised: 59 us for
  // 650 k rows, the longest kernel of a DynamicScatter call)
  __shared__ long long s_lo[8][kMaxK], s_hi[8][kMaxK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < K; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      const long long a = __shfl_xor_sync(0xffffffffu, lo[k], o), b = __shfl_xor_sync(0xffffffffu, hi[k], o);
      lo[k] = a < lo[k] ? a : lo[k];
      hi[k] = b > hi[k] ? b : hi[k];
    }
    if (lane == 0) { s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; }
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    long long a = s_lo[0][k], b = s_hi[0][k];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a = s_lo[w][k] < a ? s_lo[w][k] : a;
      b = s_hi[w][k] > b ? s_hi[w][k] : b;
    }
    atomicMin(&mm[2 * k], a);
    atomicMax(&mm[2 * k + 1], b);
  }
}

__global__ void k_minmax_init(long long *mm) {
  if (threadIdx.x < kMaxK) { mm[2 * threadIdx.x] = LLONG_MAX; mm[2 * threadIdx.x + 1] = LLONG_MIN; }
}

template <typename CT, typename KT>            // KT: uint32_t when the packed key fits 32 bits (less sort traffic), else uint64_t
__global__ void k_make_keys(const CT *__restrict__ coors, int64_t N, KeySpec sp, KT *__restrict__ keys,
                            int32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const CT *r = coors + i * sp.K;
  uint64_t key = 0;
  bool valid = true;
  const int first = (sp.mode == 2) ? 1 : 0;
  for (int k = 0; k < sp.K; ++k) {
    const long long v = (long long)r[k];
    if (sp.mode != 0 && k >= first && v < 0) valid = false;
    if (v > sp.vmax[k] && sp.oob) *sp.oob = 1ull;                 // (only valid rows matter, but any overflow would corrupt a key)
    key |= (uint64_t)(v - sp.bias[k]) << sp.shift[k];
  }
  if (sp.mode == 1) {
    key = valid ? (key | (1ull << sp.flag_shift)) : 0ull;          // invalid rows collapse into ONE group: (-1,-1,-1)
  } else if (sp.mode == 2) {
    const long long b = (long long)r[0];
    if (b < 0 || b >= sp.batch_size) key = (uint64_t)sp.batch_size << sp.shift[0];   // never selected by the per-sample loop
    else if (valid) key |= (1ull << sp.flag_shift);
    else key = (uint64_t)(b - sp.bias[0]) << sp.shift[0];         // the sample's (-1,-1,-1) group
  }
  keys[i] = (KT)key;
  idx[i] = (int32_t)i;
}

// heads of equal-key runs; which heads open a dropped group
template <typename KT>
__global__ void k_heads(const KT *__restrict__ keys, int64_t N, KeySpec sp, int32_t *__restrict__ seg_start,
                        int32_t *__restrict__ kept_head, uint8_t *__restrict__ drop_head) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const uint64_t k = (uint64_t)keys[j];
  const bool head = (j == 0) || keys[j] != keys[j - 1];
  bool drop = false;
  if (head) {
    if (sp.mode == 1) {
      drop = (j == 0);                                            // "first element is always (-1,-1,-1)" (:207)
    } else if (sp.mode == 2) {
      const int bs = sp.shift[0];
      drop = ((long long)(k >> bs) == sp.batch_size) || (j == 0) || ((k >> bs) != ((uint64_t)keys[j - 1] >> bs));   // first unique row of each sample
    }
  }
  seg_start[j] = head ? (int32_t)j : 0;
  kept_head[j] = (head && !drop) ? 1 : 0;
  drop_head[j] = drop ? 1 : 0;
}

template <typename CT, typename KT>
__global__ void k_emit(const CT *__restrict__ coors, int K, const KT *__restrict__ keys,
                       const int32_t *__restrict__ idx, int64_t N, const int32_t *__restrict__ seg_start,
                       const int32_t *__restrict__ kept_scan, const uint8_t *__restrict__ drop_head,
                       CT *__restrict__ uniq, int32_t *__restrict__ inverse, int32_t *__restrict__ counts,
                       int32_t *__restrict__ gstart) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int s = seg_start[j];
  const bool dropped = drop_head[s];
  const int og = dropped ? -1 : kept_scan[s] - 1;
  const int p = idx[j];
  inverse[p] = og;
  if (dropped) return;
  if (j == s) {
    for (int k = 0; k < K; ++k) uniq[(int64_t)og * K + k] = coors[(int64_t)p * K + k];
    gstart[og] = s;
  }
  if (j == N - 1 || keys[j + 1] != keys[j]) counts[og] = (int32_t)(j - s + 1);
}

struct MaxOpI {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

struct UqLayout {
  int64_t mm, keys_a, keys_b, idx_a, seg_start, kept, drop, cub, cub_bytes, bytes;
};

static UqLayout uq_layout(int64_t N) {
  UqLayout l;
  int64_t off = 0;
  auto take = [&](int64_t b) { int64_t o = off; off = align_up(off + b, 256); return o; };
  const int64_t n = N > 0 ? N : 1;
  l.mm = take(8 * 2 * kMaxK + 8);
  l.keys_a = take(8 * n); l.keys_b = take(8 * n);
  l.idx_a = take(4 * n);
  l.seg_start = take(4 * n); l.kept = take(4 * n); l.drop = take(n);
  size_t s1 = 0, s2 = 0, s3 = 0;
  size_t s1b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, s1, (uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, s1b, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)n);
  s1 = std::max(s1, s1b);
  cub::DeviceScan::InclusiveSum(nullptr, s2, (int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  cub::DeviceScan::InclusiveScan(nullptr, s3, (int32_t *)nullptr, (int32_t *)nullptr, MaxOpI(), (int)n);
  l.cub_bytes = (int64_t)std::max(s1, std::max(s2, s3));
  l.cub = take(l.cub_bytes);
  l.bytes = off;
  return l;
}

static int bit_length(unsigned long long v) {
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return b;
}

// keys -> stable radix sort of (key, index) -> heads -> unique rows / inverse / counts / plan
template <typename CT, typename KT>
static int sort_and_emit(const CT *coors, int64_t N, int K, const KeySpec &sp, int end_bit, CT *uniq, int32_t *inverse,
                         int32_t *counts, int32_t *order, int32_t *gstart, char *ws, const UqLayout &l,
                         cudaStream_t stream) {
  KT *keys_a = (KT *)(ws + l.keys_a), *keys_b = (KT *)(ws + l.keys_b);
  int32_t *idx_a = (int32_t *)(ws + l.idx_a);
  int32_t *seg_start = (int32_t *)(ws + l.seg_start), *kept = (int32_t *)(ws + l.kept);
  uint8_t *drop = (uint8_t *)(ws + l.drop);
  const unsigned grid = (unsigned)ceil_div(N, 256);
  k_make_keys<CT, KT><<<grid, 256, 0, stream>>>(coors, N, sp, keys_a, idx_a);
  OCC_KERNEL_OK("k_make_keys");
  size_t cb = (size_t)l.cub_bytes;
  OCC_CUDA(cub::DeviceRadixSort::SortPairs(ws + l.cub, cb, keys_a, keys_b, idx_a, order, (int)N, 0, end_bit, stream));
  count_launch(3);
  k_heads<KT><<<grid, 256, 0, stream>>>(keys_b, N, sp, seg_start, kept, drop);
  OCC_KERNEL_OK("k_heads");
  cb = (size_t)l.cub_bytes;
  OCC_CUDA(cub::DeviceScan::InclusiveScan(ws + l.cub, cb, seg_start, seg_start, MaxOpI(), (int)N, stream));
  cb = (size_t)l.cub_bytes;
  OCC_CUDA(cub::DeviceScan::InclusiveSum(ws + l.cub, cb, kept, kept, (int)N, stream));
  count_launch(4);
  k_emit<CT, KT><<<grid, 256, 0, stream>>>(coors, K, keys_b, order, N, seg_start, kept, drop, uniq, inverse, counts, gstart);
  OCC_KERNEL_OK("k_emit");
  return 0;
}

template <typename CT>
static int unique_impl(const CT *coors, int64_t N, int K, int mode, const int64_t *col_max, CT *uniq, int32_t *inverse,
                       int32_t *counts, int32_t *order, int32_t *gstart, char *ws, const UqLayout &l, int64_t *m_host,
                       cudaStream_t stream) {
  long long *mm = (long long *)(ws + l.mm);
  // Bounded call (modes 1 / 2, every bound >= 0): the caller knows the grid the coordinates come from, so the key
  // widths need no min/max pass over the rows and (mode 1) no host round trip before the sort.  A row beyond a bound
  // raises a device flag that is read with the group count; the call is then repeated unbounded.
  const bool bounded = col_max != nullptr && mode != 0;
  long long h_mm[2 * kMaxK + 1];
  CT last0 = 0;
  if (!bounded) {
    k_minmax_init<<<1, 32, 0, stream>>>(mm);
    OCC_KERNEL_OK("k_minmax_init");
    const int mm_grid = (int)std::min<int64_t>(ceil_div(N, 256 * 4), kNumSMs * 4);
    k_minmax<CT><<<mm_grid, 256, 0, stream>>>(coors, N, K, mm);
    OCC_KERNEL_OK("k_minmax");
    OCC_CUDA(cudaMemcpyAsync(h_mm, mm, 8 * 2 * K, cudaMemcpyDeviceToHost, stream));
  } else {
    for (int k = 0; k < K; ++k) {
      OCC_REQUIRE(col_max[k] >= 0, "col_max must be non-negative");
      h_mm[2 * k] = 0;
      h_mm[2 * k + 1] = (long long)col_max[k];
    }
    OCC_CUDA(cudaMemsetAsync(mm + 2 * kMaxK, 0, 8, stream));       // the out-of-bounds flag
  }
  if (mode == 2) OCC_CUDA(cudaMemcpyAsync(&last0, coors + (N - 1) * K, sizeof(CT), cudaMemcpyDeviceToHost, stream));
  if (!bounded || mode == 2) OCC_CUDA(cudaStreamSynchronize(stream));

  KeySpec sp;
  sp.K = K;
  sp.mode = mode;
  sp.batch_size = (long long)last0 + 1;          // scatter_points.py:86
  sp.oob = bounded ? (unsigned long long *)(mm + 2 * kMaxK) : nullptr;
  for (int k = 0; k < kMaxK; ++k) sp.vmax[k] = LLONG_MAX;
  if (bounded)
    for (int k = (mode == 2 ? 1 : 0); k < K; ++k) sp.vmax[k] = (long long)col_max[k];   // (the batch column is bounded by batch_size)
  int bits[kMaxK];
  for (int k = 0; k < K; ++k) {
    long long lo = h_mm[2 * k], hi = h_mm[2 * k + 1];
    if (mode == 1 || (mode == 2 && k > 0)) lo = 0;                 // valid rows are >= 0; invalid rows do not use the field
    if (mode == 2 && k == 0) { lo = 0; hi = std::max<long long>(sp.batch_size, 0); }   // +1 value: out-of-range sentinel
    if (hi < lo) hi = lo;
    sp.bias[k] = lo;
    bits[k] = bit_length((unsigned long long)(hi - lo));
  }
  int total = 0;
  if (mode == 2) {
    // layout (msb..lsb): batch | valid flag | col1 .. colK-1
    for (int k = K - 1; k >= 1; --k) { sp.shift[k] = total; total += bits[k]; }
    sp.flag_shift = total; total += 1;
    sp.shift[0] = total; total += bits[0];
  } else {
    for (int k = K - 1; k >= 0; --k) { sp.shift[k] = total; total += bits[k]; }
    sp.flag_shift = total;
    if (mode == 1) total += 1;
  }
  OCC_REQUIRE(total <= 63, "coordinate ranges need more than 63 key bits");
  const int end_bit = std::max(total, 1);

  int32_t *kept = (int32_t *)(ws + l.kept);
  if (total <= 32) {
    if (sort_and_emit<CT, uint32_t>(coors, N, K, sp, end_bit, uniq, inverse, counts, order, gstart, ws, l, stream)) return 1;
  } else {
    if (sort_and_emit<CT, uint64_t>(coors, N, K, sp, end_bit, uniq, inverse, counts, order, gstart, ws, l, stream)) return 1;
  }
  int32_t m32 = 0;
  unsigned long long oob = 0;
  OCC_CUDA(cudaMemcpyAsync(&m32, kept + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
  if (bounded) OCC_CUDA(cudaMemcpyAsync(&oob, mm + 2 * kMaxK, 8, cudaMemcpyDeviceToHost, stream));
  OCC_CUDA(cudaStreamSynchronize(stream));
  if (oob)                                         // a coordinate beyond the caller's bound: the keys are not trustworthy
    return unique_impl<CT>(coors, N, K, mode, nullptr, uniq, inverse, counts, order, gstart, ws, l, m_host, stream);
  *m_host = m32;
  return 0;
}

// --------------------------------------------------------------------------- plan from inverse
__global__ void k_inv_keys(const int32_t *__restrict__ inverse, int64_t N, uint32_t *__restrict__ keys,
                           int32_t *__restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  keys[i] = (uint32_t)(inverse[i] + 1);
  idx[i] = (int32_t)i;
}

__global__ void k_inv_plan(const uint32_t *__restrict__ keys, int64_t N, int32_t *__restrict__ gstart,
                           int32_t *__restrict__ counts) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const uint32_t k = keys[j];
  if (k == 0) return;                                             // inverse == -1
  if (j == 0 || keys[j - 1] != k) gstart[k - 1] = (int32_t)j;
  if (counts) atomicAdd(&counts[k - 1], 1);
}

struct PlLayout {
  int64_t keys_a, keys_b, idx_a, cub, cub_bytes, bytes;
};
static PlLayout pl_layout(int64_t N) {
  PlLayout l;
  int64_t off = 0;
  auto take = [&](int64_t b) { int64_t o = off; off = align_up(off + b, 256); return o; };
  const int64_t n = N > 0 ? N : 1;
  l.keys_a = take(4 * n); l.keys_b = take(4 * n); l.idx_a = take(4 * n);
  size_t s1 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, s1, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)n);
  l.cub_bytes = (int64_t)s1;
  l.cub = take(l.cub_bytes);
  l.bytes = off;
  return l;
}

// --------------------------------------------------------------------------- segmented reductions
// WIDTH lanes cooperate on one voxel; lane = channel (+ WIDTH strides when C > WIDTH).
// exact IEEE add / divide in either precision (never contracted or reordered)
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <int WIDTH, typename F>
__global__ void __launch_bounds__(256)
k_segment_reduce(const F *__restrict__ feats, int C, const int32_t *__restrict__ order,
                 const int32_t *__restrict__ gstart, const int32_t *__restrict__ counts, int64_t M, int reduce,
                 int64_t N, F *__restrict__ out, int32_t *__restrict__ argmax) {
  const int64_t g = ((int64_t)blockIdx.x * 256 + threadIdx.x) / WIDTH;
  const int lane = threadIdx.x % WIDTH;
  if (g >= M) return;
  const int s = gstart[g], n = counts[g];
  for (int c = lane; c < C; c += WIDTH) {
    F acc = (reduce == OCCB200_MAX) ? (F)-INFINITY : (F)0;
    int32_t arg = (int32_t)N;
    for (int j = 0; j < n; ++j) {
      const int p = __ldg(order + s + j);
      const F v = __ldg(feats + (int64_t)p * C + c);
      if (reduce == OCCB200_MAX) {
        if (v > acc) { acc = v; arg = p; }       // ascending p: the first attaining index is kept
      } else {
        acc = add_rn(acc, v);
      }
    }
    if (reduce == OCCB200_MEAN) acc = div_rn(acc, (F)(n < 1 ? 1 : n));
    out[g * C + c] = acc;
    if (argmax) argmax[g * C + c] = arg;
  }
}

template <typename F>
__global__ void k_bwd_add(F *__restrict__ grad_feats, const F *__restrict__ grad_reduced,
                          const int32_t *__restrict__ inverse, const int32_t *__restrict__ counts, int64_t N, int C,
                          int reduce) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * C) return;
  const int64_t i = e / C;
  const int c = (int)(e % C);
  const int v = inverse[i];
  F g = (F)0;
  if (v >= 0) {
    g = grad_reduced[(int64_t)v * C + c];
    if (reduce == OCCB200_MEAN) g = div_rn(g, (F)counts[v]);   // :126-129
  }
  grad_feats[e] = g;
}

template <typename F>
__global__ void k_bwd_argmin(const F *__restrict__ feats, const F *__restrict__ reduced,
                             const int32_t *__restrict__ inverse, int64_t N, int C, int32_t *__restrict__ from) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * C) return;
  const int64_t i = e / C;
  const int c = (int)(e % C);
  const int v = inverse[i];
  if (v < 0) return;
  if (feats[e] == reduced[(int64_t)v * C + c]) atomicMin(&from[(int64_t)v * C + c], (int32_t)i);   // :152-157
}

__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) p[e] = v;
}

template <typename F>
__global__ void k_bwd_max_scatter(F *__restrict__ grad_feats, const F *__restrict__ grad_reduced,
                                  const int32_t *__restrict__ from, int64_t M, int C, int64_t N) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * C) return;
  const int c = (int)(e % C);
  const int p = from[e];
  if (p >= 0 && p < N) grad_feats[(int64_t)p * C + c] = grad_reduced[e];   // :171-176
}

}  // namespace occb200

using namespace occb200;

extern "C" int64_t occb200_unique_workspace_bytes(int64_t N, int K) {
  (void)K;
  return uq_layout(N).bytes;
}

extern "C" int occb200_unique_rows_bounded(const void *coors, int coor_dtype, int64_t N, int K, int mode,
                                           const int64_t *col_max, void *uniq, int32_t *inverse, int32_t *counts,
                                           int32_t *order, int32_t *gstart, void *workspace, int64_t workspace_bytes,
                                           int64_t *m_host, void *stream_);

extern "C" int occb200_unique_rows(const void *coors, int coor_dtype, int64_t N, int K, int mode, void *uniq,
                                   int32_t *inverse, int32_t *counts, int32_t *order, int32_t *gstart,
                                   void *workspace, int64_t workspace_bytes, int64_t *m_host, void *stream_) {
  return occb200_unique_rows_bounded(coors, coor_dtype, N, K, mode, nullptr, uniq, inverse, counts, order, gstart,
                                     workspace, workspace_bytes, m_host, stream_);
}

extern "C" int occb200_unique_rows_bounded(const void *coors, int coor_dtype, int64_t N, int K, int mode,
                                           const int64_t *col_max, void *uniq, int32_t *inverse, int32_t *counts,
                                           int32_t *order, int32_t *gstart, void *workspace, int64_t workspace_bytes,
                                           int64_t *m_host, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(m_host != nullptr, "m_host is NULL");
  *m_host = 0;
  OCC_REQUIRE(N >= 0 && N < (1ll << 31), "N must fit int32");
  OCC_REQUIRE(K >= 1 && K <= kMaxK, "1 <= K <= 8");
  OCC_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  OCC_REQUIRE(mode != 2 || K >= 2, "mode 2 needs a batch column");
  OCC_REQUIRE(coor_dtype == 0 || coor_dtype == 1, "coor_dtype must be 0 (int32) or 1 (int64)");
  if (N == 0) return 0;
  const UqLayout l = uq_layout(N);
  OCC_REQUIRE(workspace != nullptr && workspace_bytes >= l.bytes, "workspace too small");
  if (coor_dtype == 0)
    return unique_impl<int32_t>((const int32_t *)coors, N, K, mode, col_max, (int32_t *)uniq, inverse, counts, order,
                                gstart, (char *)workspace, l, m_host, stream);
  return unique_impl<int64_t>((const int64_t *)coors, N, K, mode, col_max, (int64_t *)uniq, inverse, counts, order, gstart,
                              (char *)workspace, l, m_host, stream);
}

extern "C" int64_t occb200_plan_workspace_bytes(int64_t N) { return pl_layout(N).bytes; }

extern "C" int occb200_plan_from_inverse(const int32_t *inverse, int64_t N, int64_t M, int32_t *order, int32_t *gstart,
                                         int32_t *counts, void *workspace, int64_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(N >= 0 && N < (1ll << 31) && M >= 0 && M < (1ll << 31) - 1, "N, M must fit int32");
  if (M > 0) {
    OCC_CUDA(cudaMemsetAsync(gstart, 0, 4 * M, stream));
    if (counts) OCC_CUDA(cudaMemsetAsync(counts, 0, 4 * M, stream));
  }
  if (N == 0) return 0;
  const PlLayout l = pl_layout(N);
  OCC_REQUIRE(workspace != nullptr && workspace_bytes >= l.bytes, "workspace too small");
  char *ws = (char *)workspace;
  uint32_t *keys_a = (uint32_t *)(ws + l.keys_a), *keys_b = (uint32_t *)(ws + l.keys_b);
  int32_t *idx_a = (int32_t *)(ws + l.idx_a);
  const unsigned grid = (unsigned)ceil_div(N, 256);
  k_inv_keys<<<grid, 256, 0, stream>>>(inverse, N, keys_a, idx_a);
  OCC_KERNEL_OK("k_inv_keys");
  size_t cb = (size_t)l.cub_bytes;
  const int end_bit = std::max(bit_length((unsigned long long)M), 1);
  OCC_CUDA(cub::DeviceRadixSort::SortPairs(ws + l.cub, cb, keys_a, keys_b, idx_a, order, (int)N, 0, end_bit, stream));
  count_launch(3);
  k_inv_plan<<<grid, 256, 0, stream>>>(keys_b, N, gstart, counts);
  OCC_KERNEL_OK("k_inv_plan");
  return 0;
}

template <typename F>
static int segment_reduce_impl(const F *feats, int64_t N, int C, const int32_t *order, const int32_t *gstart,
                               const int32_t *counts, int64_t M, int reduce, F *out, int32_t *argmax,
                               cudaStream_t stream) {
  OCC_REQUIRE(C >= 1 && M >= 0 && N >= 0, "bad sizes");
  OCC_REQUIRE(reduce >= 0 && reduce <= 2, "do not support reduce type");
  if (M == 0) return 0;
#define OCC_SR(W)                                                                                         \
  k_segment_reduce<W, F><<<(unsigned)ceil_div(M * W, 256), 256, 0, stream>>>(feats, C, order, gstart, counts, M, \
                                                                             reduce, N, out, argmax)
  if (C <= 4) OCC_SR(4);
  else if (C <= 8) OCC_SR(8);
  else if (C <= 16) OCC_SR(16);
  else OCC_SR(32);
#undef OCC_SR
  OCC_KERNEL_OK("k_segment_reduce");
  return 0;
}

template <typename F>
static int segment_reduce_backward_impl(F *grad_feats, const F *grad_reduced, const F *feats, const F *reduced_feats,
                                        const int32_t *inverse, const int32_t *counts, const int32_t *argmax, int64_t N,
                                        int64_t M, int C, int reduce, cudaStream_t stream) {
  OCC_REQUIRE(C >= 1 && M >= 0 && N >= 0, "bad sizes");
  OCC_REQUIRE(reduce >= 0 && reduce <= 2, "do not support reduce type");
  if (N == 0) return 0;
  if (M == 0 || reduce == OCCB200_MAX) OCC_CUDA(cudaMemsetAsync(grad_feats, 0, sizeof(F) * N * C, stream));   // :254
  if (M == 0) return 0;
  if (reduce != OCCB200_MAX) {
    k_bwd_add<F><<<(unsigned)ceil_div(N * C, 256), 256, 0, stream>>>(grad_feats, grad_reduced, inverse, counts, N, C, reduce);
    OCC_KERNEL_OK("k_bwd_add");
    return 0;
  }
  int32_t *from = nullptr;
  if (!argmax) {
    OCC_REQUIRE(feats && reduced_feats, "max backward needs feats and reduced_feats (or argmax)");
    OCC_CUDA(cudaMallocAsync((void **)&from, 4 * M * C, stream));
    k_fill_i32<<<(unsigned)ceil_div(M * C, 256), 256, 0, stream>>>(from, M * C, (int32_t)N);
    OCC_KERNEL_OK("k_fill_i32");
    k_bwd_argmin<F><<<(unsigned)ceil_div(N * C, 256), 256, 0, stream>>>(feats, reduced_feats, inverse, N, C, from);
    OCC_KERNEL_OK("k_bwd_argmin");
    argmax = from;
  }
  k_bwd_max_scatter<F><<<(unsigned)ceil_div(M * C, 256), 256, 0, stream>>>(grad_feats, grad_reduced, argmax, M, C, N);
  OCC_KERNEL_OK("k_bwd_max_scatter");
  if (from) OCC_CUDA(cudaFreeAsync(from, stream));
  return 0;
}

extern "C" int occb200_segment_reduce(const float *feats, int64_t N, int C, const int32_t *order, const int32_t *gstart,
                                      const int32_t *counts, int64_t M, int reduce, float *out, int32_t *argmax,
                                      void *stream) {
  return segment_reduce_impl<float>(feats, N, C, order, gstart, counts, M, reduce, out, argmax, (cudaStream_t)stream);
}

extern "C" int occb200_segment_reduce_f64(const double *feats, int64_t N, int C, const int32_t *order,
                                          const int32_t *gstart, const int32_t *counts, int64_t M, int reduce,
                                          double *out, int32_t *argmax, void *stream) {
  return segment_reduce_impl<double>(feats, N, C, order, gstart, counts, M, reduce, out, argmax, (cudaStream_t)stream);
}

extern "C" int occb200_segment_reduce_backward(float *grad_feats, const float *grad_reduced, const float *feats,
                                               const float *reduced_feats, const int32_t *inverse,
                                               const int32_t *counts, const int32_t *argmax, int64_t N, int64_t M,
                                               int C, int reduce, void *stream) {
  return segment_reduce_backward_impl<float>(grad_feats, grad_reduced, feats, reduced_feats, inverse, counts, argmax, N,
                                             M, C, reduce, (cudaStream_t)stream);
}

extern "C" int occb200_segment_reduce_backward_f64(double *grad_feats, const double *grad_reduced, const double *feats,
                                                   const double *reduced_feats, const int32_t *inverse,
                                                   const int32_t *counts, const int32_t *argmax, int64_t N, int64_t M,
                                                   int C, int reduce, void *stream) {
  return segment_reduce_backward_impl<double>(grad_feats, grad_reduced, feats, reduced_feats, inverse, counts, argmax,
                                              N, M, C, reduce, (cudaStream_t)stream);
}
