// annotate.cu -- batched tracklet annotation: crop -> box frame -> voxelise -> visibility.
//
// Reference path: tools/occ/occ_annotate.py get_local_point_list :91-138 and
// OccAnnotator.annotate_trk :344-568 (normative step list: SURVEY.md Appendix A).
//
// Kernels (all on the caller's stream, no host synchronisation):
//   k_frame_inbox     one CTA per tracklet-frame: does any candidate point fall in the box?   (A1)
//   k_tracklet_setup  one thread per tracklet: box size = max over kept frames, dims, bounds  (A2/A3)
//   k_scan_chunks     one CTA: exclusive scan of per-tracklet work chunks -> work list
//   k_frame_voxelize  one CTA per tracklet-frame: in-box -> box frame -> quantise -> bitset   (A2/A3)
//   k_visibility      persistent CTAs over 256-voxel chunks: the range-image "ray-cast"       (A4/A5)
#include <math.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "geom.cuh"

namespace occb200 {

constexpr int kChunk = 256;         // voxels per work item == threads per visibility CTA
constexpr int kFrameThreads = 256;
constexpr int kSmemBitWords = 8192; // 32 KB: grids up to 262 144 voxels keep their bitset in shared memory

struct TrkGrid {
  int32_t dims[3];
  int32_t status;
  float mb[3];      // min bound of the canonical box (f32)
  int32_t flags;    // bit0: some point survived the q<dims filter, bit1: index error
  int64_t V;
  int64_t bits_off; // word offset of the occupancy bitset
  int32_t nchunks;
  int32_t B;
};

struct Workspace {
  TrkGrid *grids;        // [T]
  int32_t *frame_kept;   // [F]
  int32_t *frame_trk;    // [F]
  int64_t *chunk_off;    // [T+1]
  unsigned long long *counter;   // work-queue head
  uint32_t *bits;        // occupancy bitsets
  int64_t bits_words;
};

static int64_t ws_layout(int32_t T, int64_t F, int64_t total, char *base, Workspace *w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  int64_t o_grid = take(sizeof(TrkGrid) * (int64_t)T);
  int64_t o_kept = take(4 * F);
  int64_t o_ftrk = take(4 * F);
  int64_t o_choff = take(8 * ((int64_t)T + 1));
  int64_t o_cnt = take(8);
  int64_t words = total / 32 + T + 1;
  int64_t o_bits = take(4 * words);
  if (w) {
    w->grids = (TrkGrid *)(base + o_grid);
    w->frame_kept = (int32_t *)(base + o_kept);
    w->frame_trk = (int32_t *)(base + o_ftrk);
    w->chunk_off = (int64_t *)(base + o_choff);
    w->counter = (unsigned long long *)(base + o_cnt);
    w->bits = (uint32_t *)(base + o_bits);
    w->bits_words = words;
  }
  return off;
}

// ---------------------------------------------------------------------------------------------
// Optional per-kernel timing (bench.py's roofline): CUDA events recorded around each kernel of the
// pipeline on the caller's stream; durations are summed per kernel when the profile is read.
enum { kProfInbox = 0, kProfSetup, kProfScan, kProfVoxelize, kProfVisibility, kProfKinds };
struct ProfEntry { int kind; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;
static std::mutex g_prof_mu;

struct ProfScope {
  cudaStream_t stream;
  ProfEntry e;
  bool on;
  ProfScope(int kind, cudaStream_t s) : stream(s), on(g_prof_on) {
    if (!on) return;
    e.kind = kind;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(e.b, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(e);
  }
};

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFrameThreads)
k_frame_inbox(const occb200_pose_t *__restrict__ poses, const float *__restrict__ points, int stride,
              const int64_t *__restrict__ frame_pt_off, int32_t *__restrict__ frame_kept) {
  const int64_t f = blockIdx.x;
  const occb200_pose_t &ps = poses[f];
  const BoxTest bt = make_box_test(ps.box, ps.cos_pib, ps.sin_pib);
  const int64_t n0 = frame_pt_off[f], n1 = frame_pt_off[f + 1];
  int any = 0;
  for (int64_t j = n0 + threadIdx.x; j < n1; j += kFrameThreads) {
    const float *p = points + j * stride;
    any |= pt_in_box(bt, ld_stream(p), ld_stream(p + 1), ld_stream(p + 2));
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) frame_kept[f] = any;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_tracklet_setup(int T, const int64_t *__restrict__ trk_frame_off,
                                 const occb200_pose_t *__restrict__ poses,
                                 const int32_t *__restrict__ frame_kept, const int64_t *__restrict__ label_off,
                                 float vsf, TrkGrid *__restrict__ grids, int32_t *__restrict__ frame_trk,
                                 int32_t *__restrict__ dims_out, float *__restrict__ sizes_out,
                                 int32_t *__restrict__ status_out, int64_t *__restrict__ n_unknown,
                                 int64_t *__restrict__ n_steps) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int64_t f0 = trk_frame_off[t], f1 = trk_frame_off[t + 1];
  const int B = (int)(f1 - f0);
  TrkGrid g;
  g.B = B;
  g.flags = 0;
  g.nchunks = 0;
  g.V = 0;
  g.bits_off = label_off[t] / 32 + t;
  g.dims[0] = g.dims[1] = g.dims[2] = 0;
  g.mb[0] = g.mb[1] = g.mb[2] = 0.f;
  float sz[3] = {-INFINITY, -INFINITY, -INFINITY};
  int kept = 0;
  for (int64_t f = f0; f < f1; ++f) {
    frame_trk[f] = t;
    if (frame_kept[f]) {                       // occ_annotate.py:111-112, :132-133 (box_mode="max")
      ++kept;
      sz[0] = fmaxf(sz[0], poses[f].box[3]);
      sz[1] = fmaxf(sz[1], poses[f].box[4]);
      sz[2] = fmaxf(sz[2], poses[f].box[5]);
    }
  }
  if (B < 10) {
    g.status = OCCB200_SKIP_SHORT;             // :344
  } else if (kept == 0) {
    g.status = OCCB200_NO_POINTS;              // :129
  } else {
    g.status = OCCB200_OK;
    for (int k = 0; k < 3; ++k) g.dims[k] = (int)ceilf(__fdiv_rn(sz[k], vsf));   // :414-416
    g.mb[0] = __fmul_rn(sz[0], -0.5f);         // min over corners of [0,0,0,w,l,h,0] (:422-423)
    g.mb[1] = __fmul_rn(sz[1], -0.5f);
    g.mb[2] = __fmul_rn(sz[2], 0.0f);
    g.V = (int64_t)g.dims[0] * g.dims[1] * g.dims[2];
    if (g.V > label_off[t + 1] - label_off[t] || g.V <= 0) {
      g.status = -1;                           // caller's slot too small: reported, nothing written
      g.V = 0;
    }
    g.nchunks = (int)((g.V + kChunk - 1) / kChunk);
  }
  grids[t] = g;
  for (int k = 0; k < 3; ++k) {
    dims_out[3 * t + k] = g.dims[k];
    sizes_out[3 * t + k] = (g.status == OCCB200_OK) ? sz[k] : 0.f;
  }
  status_out[t] = g.status;                    // refined by k_visibility (flags)
  n_unknown[t] = 0;
  if (n_steps) n_steps[t] = 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_scan_chunks(int T, const TrkGrid *__restrict__ grids, int64_t *__restrict__ chunk_off,
              unsigned long long *__restrict__ counter) {
  __shared__ int64_t s_part[1024];
  const int tid = threadIdx.x;
  const int per = (T + 1023) / 1024;
  const int a = min(tid * per, T), b = min(a + per, T);
  int64_t sum = 0;
  for (int t = a; t < b; ++t) sum += grids[t].nchunks;
  s_part[tid] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {          // Hillis-Steele inclusive scan
    int64_t v = (tid >= d) ? s_part[tid - d] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int64_t run = s_part[tid] - sum;
  for (int t = a; t < b; ++t) {
    chunk_off[t] = run;
    run += grids[t].nchunks;
  }
  if (tid == 1023) chunk_off[T] = s_part[1023];
  if (tid == 0) *counter = 0ull;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFrameThreads)
k_frame_voxelize(const occb200_pose_t *__restrict__ poses, const float *__restrict__ points, int stride,
                 const int64_t *__restrict__ frame_pt_off, const int32_t *__restrict__ frame_kept,
                 const int32_t *__restrict__ frame_trk, TrkGrid *__restrict__ grids,
                 uint32_t *__restrict__ bits, float vsf) {
  __shared__ uint32_t s_bits[kSmemBitWords];
  __shared__ int s_flags;
  const int64_t f = blockIdx.x;
  if (!frame_kept[f]) return;                   // frame contributes no points (:111-112)
  const int t = frame_trk[f];
  const TrkGrid g = grids[t];
  if (g.status != OCCB200_OK) return;
  const int words = (int)((g.V + 31) / 32);
  const bool use_smem = words <= kSmemBitWords;
  uint32_t *gbits = bits + g.bits_off;
  if (use_smem)
    for (int w = threadIdx.x; w < words; w += kFrameThreads) s_bits[w] = 0u;
  if (threadIdx.x == 0) s_flags = 0;
  __syncthreads();

  const occb200_pose_t &ps = poses[f];
  const BoxTest bt = make_box_test(ps.box, ps.cos_pib, ps.sin_pib);
  const float ox = ps.box[0], oy = ps.box[1], oz = ps.box[2];
  const float c = ps.cos_m, s = ps.sin_m;       // torch f32 cos/sin(-yaw)
  const float dX = (float)g.dims[0], dY = (float)g.dims[1], dZ = (float)g.dims[2];
  const int64_t n0 = frame_pt_off[f], n1 = frame_pt_off[f + 1];
  const int lane = threadIdx.x & 31;
  int flags = 0;
  for (int64_t base = n0; base < n1; base += kFrameThreads) {   // warp-uniform trip count
    const int64_t j = base + threadIdx.x;
    int word = -1;
    uint32_t bit = 0u;
    if (j < n1) {
      const float *p = points + j * stride;
      const float x = ld_stream(p), y = ld_stream(p + 1), z = ld_stream(p + 2);
      if (pt_in_box(bt, x, y, z)) {
        // local = (p + (-origin)) @ [[c,-s,0],[s,c,0],[0,0,1]]  (:117-122, lidar_box3d.py:165-184):
        // sgemm accumulates k = 0,1,2 as an FMA chain; the k=2 terms are exact no-ops.
        const float tx = __fadd_rn(x, -ox), ty = __fadd_rn(y, -oy), tz = __fadd_rn(z, -oz);
        const float lx = __fmaf_rn(ty, s, __fmul_rn(tx, c));
        const float ly = __fmaf_rn(ty, c, __fmul_rn(tx, -s));
        const float lz = tz;
        // q = floor((local - min_bound) / vs)  (:425)
        float qx = floorf(__fdiv_rn(__fsub_rn(lx, g.mb[0]), vsf));
        float qy = floorf(__fdiv_rn(__fsub_rn(ly, g.mb[1]), vsf));
        float qz = floorf(__fdiv_rn(__fsub_rn(lz, g.mb[2]), vsf));
        if (qx < dX && qy < dY && qz < dZ) {    // only the upper bound is filtered (:430-431)
          flags |= 1;
          if (qx < 0.f) qx += dX;               // PyTorch negative-index wrap (:436)
          if (qy < 0.f) qy += dY;
          if (qz < 0.f) qz += dZ;
          if (qx < 0.f || qy < 0.f || qz < 0.f) {
            flags |= 2;                         // IndexError in the reference
          } else {
            const int64_t idx = ((int64_t)qx * g.dims[1] + (int64_t)qy) * g.dims[2] + (int64_t)qz;
            word = (int)(idx >> 5);
            bit = 1u << (idx & 31);
          }
        }
      }
    }
    // warp-level dedup: lanes that hit the same bitset word merge their bits, one atomic per word
    const unsigned peers = __match_any_sync(0xffffffffu, word);
    const uint32_t merged = __reduce_or_sync(peers, bit);
    if (word >= 0 && lane == __ffs(peers) - 1) {
      if (use_smem) {
        if ((s_bits[word] & merged) != merged) atomicOr(&s_bits[word], merged);
      } else {
        atomicOr(&gbits[word], merged);
      }
    }
  }
  if (flags) atomicOr(&s_flags, flags);
  __syncthreads();
  if (use_smem)
    for (int w = threadIdx.x; w < words; w += kFrameThreads) {
      const uint32_t v = s_bits[w];
      if (v) atomicOr(&gbits[w], v);
    }
  if (threadIdx.x == 0 && s_flags) atomicOr(&grids[t].flags, s_flags);
}

// ---------------------------------------------------------------------------------------------
// Visibility: label every voxel without a point as free (2) if, for ANY frame and ANY LiDAR, the
// range image holds a return at least as far as the voxel centre along the pixel the centre
// projects to; otherwise unknown (0).  Occupied voxels are 1.  (occ_annotate.py:466-563)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ SensorView load_sensor(const occb200_sensor_t *__restrict__ sn,
                                                  const float *__restrict__ incl_pool) {
  SensorView s;
#pragma unroll
  for (int k = 0; k < 12; ++k) s.v[k] = (double)__ldg(&sn->v2l[k]);
  s.azc = (double)__ldg(&sn->azc);
  s.H = __ldg(&sn->H);
  s.W = __ldg(&sn->W);
  s.mono = __ldg(&sn->incl_mono);
  s.incl = incl_pool + __ldg(&sn->incl_off);
  return s;
}

__global__ void __launch_bounds__(kChunk)
k_visibility_f64(int T, int L, const int64_t *__restrict__ trk_frame_off,
                 const occb200_pose_t *__restrict__ poses, const int32_t *__restrict__ frame_sf,
                 const occb200_sensor_t *__restrict__ sensors, const float *__restrict__ incl_pool,
                 const float *__restrict__ ri_pool, double vs, const int64_t *__restrict__ label_off,
                 const TrkGrid *__restrict__ grids, const int64_t *__restrict__ chunk_off,
                 unsigned long long *__restrict__ counter, const uint32_t *__restrict__ bits,
                 int32_t *__restrict__ labels, int32_t *__restrict__ status_out,
                 int64_t *__restrict__ n_unknown, int64_t *__restrict__ n_steps) {
  __shared__ long long s_item;
  const int lane = threadIdx.x & 31;
  const long long total = chunk_off[T];
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = (long long)atomicAdd(counter, 1ull);
    __syncthreads();
    const long long item = s_item;
    if (item >= total) break;
    int lo = 0, hi = T;                          // last t with chunk_off[t] <= item
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (chunk_off[mid] <= item) lo = mid; else hi = mid;
    }
    const int t = lo;
    const TrkGrid g = grids[t];
    const int chunk = (int)(item - chunk_off[t]);
    // status refinement once per tracklet (flags are final: k_frame_voxelize has completed)
    int status = g.status;
    if (status == OCCB200_OK) {
      if (g.flags & 2) status = OCCB200_INDEX_ERROR;
      else if (!(g.flags & 1)) status = OCCB200_EMPTY_AFTER_FILTER;
    }
    if (chunk == 0 && threadIdx.x == 0) status_out[t] = status;
    if (status != OCCB200_OK) continue;          // the reference produces no output here

    const int64_t f = (int64_t)chunk * kChunk + threadIdx.x;
    const bool active = f < g.V;
    bool occupied = false;
    if (active) occupied = (bits[g.bits_off + (f >> 5)] >> (f & 31)) & 1u;
    const bool need = active && !occupied;
    bool is_free = false;
    long long steps = 0;
    if (__any_sync(0xffffffffu, need)) {
      const int YZ = g.dims[1] * g.dims[2];
      const int x = (int)(f / YZ), y = (int)((f / g.dims[2]) % g.dims[1]), z = (int)(f % g.dims[2]);
      // centre = coord.f64 * vs + min_bound + vs/2, left to right (:467-471)
      const double cx = __dadd_rn(__dadd_rn(__dmul_rn((double)x, vs), (double)g.mb[0]), vs / 2);
      const double cy = __dadd_rn(__dadd_rn(__dmul_rn((double)y, vs), (double)g.mb[1]), vs / 2);
      const double cz = __dadd_rn(__dadd_rn(__dmul_rn((double)z, vs), (double)g.mb[2]), vs / 2);
      const int64_t f0 = trk_frame_off[t];
      for (int c = 0; c < L; ++c) {              // LiDARs (:525), OR-ed (:552-556)
        for (int i = 0; i < g.B; ++i) {          // frames (:479), OR-ed (:550)
          if (__all_sync(0xffffffffu, !need || is_free)) break;   // warp-uniform early exit
          if (need && !is_free) {
            const occb200_pose_t &ps = poses[f0 + i];
            const double rc = (double)__ldg(&ps.cos_p), rs = (double)__ldg(&ps.sin_p);   // :490-496
            // ego = centre @ [[c,-s,0],[s,c,0],[0,0,1]] + origin (:497-498); z row is exact
            const double ex = __dadd_rn(__fma_rn(cy, rs, __dmul_rn(cx, rc)), (double)__ldg(&ps.box[0]));
            const double ey = __dadd_rn(__fma_rn(cy, rc, __dmul_rn(cx, -rs)), (double)__ldg(&ps.box[1]));
            const double ez = __dadd_rn(cz, (double)__ldg(&ps.box[2]));
            const occb200_sensor_t *sn = sensors + (int64_t)__ldg(frame_sf + f0 + i) * L + c;
            const SensorView sv = load_sensor(sn, incl_pool);
            int row, col;
            double rng;
            project_exact(ex, ey, ez, sv, row, col, rng);
            if (col < 0) col += sv.W;            // negative index wraps (:543)
            const float ri = __ldg(ri_pool + __ldg(&sn->ri_off) + (int64_t)row * sv.W + col);
            is_free = (double)ri >= rng;         // :547
            ++steps;
          }
        }
        if (__all_sync(0xffffffffu, !need || is_free)) break;
      }
    }
    if (active) labels[label_off[t] + f] = occupied ? 1 : (is_free ? 2 : 0);   // :558-563
    const unsigned nmask = __ballot_sync(0xffffffffu, need);
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0) {
      if (nmask) atomicAdd((unsigned long long *)&n_unknown[t], (unsigned long long)__popc(nmask));
      if (n_steps && steps) atomicAdd((unsigned long long *)&n_steps[t], (unsigned long long)steps);
    }
  }
}

// standalone operator: the reference's point_cloud_to_range_image_idx
__global__ void k_project_points(const double *__restrict__ points, int B, int64_t N,
                                 const float *__restrict__ v2l, const float *__restrict__ azc,
                                 const float *__restrict__ incl, int H, int W, int mono,
                                 int64_t *__restrict__ ri_idx, double *__restrict__ ri_range) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= N) return;
  SensorView s;
  for (int k = 0; k < 12; ++k) s.v[k] = (double)v2l[12 * b + k];
  s.azc = (double)azc[b];
  s.H = H;
  s.W = W;
  s.mono = mono;
  s.incl = incl + (int64_t)b * H;
  const double *p = points + 3 * ((int64_t)b * N + i);
  int row, col;
  double rng;
  project_exact(p[0], p[1], p[2], s, row, col, rng);
  ri_idx[2 * ((int64_t)b * N + i)] = row;
  ri_idx[2 * ((int64_t)b * N + i) + 1] = col;
  ri_range[(int64_t)b * N + i] = rng;
}

}  // namespace occb200

using namespace occb200;

extern "C" int64_t occb200_annotate_workspace_bytes(int32_t T, int64_t F, int64_t total_label_slots) {
  return ws_layout(T, F, total_label_slots, nullptr, nullptr);
}

extern "C" int occb200_annotate_batch(const occb200_annotate_args_t *a, int64_t total, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(a != nullptr, "args is NULL");
  OCC_REQUIRE(a->T >= 0 && a->F >= 0 && a->L >= 1, "bad T/F/L");
  OCC_REQUIRE(a->point_stride >= 3, "point_stride must be >= 3");
  OCC_REQUIRE(a->voxel_size > 0, "voxel_size must be positive");
  if (a->T == 0) return 0;
  Workspace w;
  const int64_t need = ws_layout(a->T, a->F, total, (char *)a->workspace, &w);
  OCC_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= need, "workspace too small");
  const float vsf = (float)a->voxel_size;
  OCC_CUDA(cudaMemsetAsync(w.bits, 0, 4 * w.bits_words, stream));
  if (a->F > 0) {
    ProfScope ps(kProfInbox, stream);
    k_frame_inbox<<<(unsigned)a->F, kFrameThreads, 0, stream>>>(a->poses, a->points, a->point_stride,
                                                                 a->frame_pt_off, w.frame_kept);
    OCC_KERNEL_OK("k_frame_inbox");
  }
  {
    ProfScope ps(kProfSetup, stream);
    k_tracklet_setup<<<(unsigned)ceil_div(a->T, 128), 128, 0, stream>>>(
        a->T, a->trk_frame_off, a->poses, w.frame_kept, a->label_off, vsf, w.grids, w.frame_trk, a->dims,
        a->sizes, a->status, a->n_unknown, a->n_steps);
    OCC_KERNEL_OK("k_tracklet_setup");
  }
  {
    ProfScope ps(kProfScan, stream);
    k_scan_chunks<<<1, 1024, 0, stream>>>(a->T, w.grids, w.chunk_off, w.counter);
    OCC_KERNEL_OK("k_scan_chunks");
  }
  if (a->F > 0) {
    ProfScope ps(kProfVoxelize, stream);
    k_frame_voxelize<<<(unsigned)a->F, kFrameThreads, 0, stream>>>(
        a->poses, a->points, a->point_stride, a->frame_pt_off, w.frame_kept, w.frame_trk, w.grids, w.bits, vsf);
    OCC_KERNEL_OK("k_frame_voxelize");
  }
  const int64_t max_items = ceil_div(total, kChunk) + a->T;
  const int grid = (int)std::min<int64_t>(max_items, (int64_t)kNumSMs * 8);
  ProfScope ps(kProfVisibility, stream);
  k_visibility_f64<<<grid, kChunk, 0, stream>>>(a->T, a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors,
                                                a->incl_pool, a->ri_pool, a->voxel_size, a->label_off, w.grids,
                                                w.chunk_off, w.counter, w.bits, a->labels, a->status,
                                                a->n_unknown, a->n_steps);
  OCC_KERNEL_OK("k_visibility_f64");
  return 0;
}

extern "C" void occb200_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
}

extern "C" int occb200_profile_read(double *ms_per_kind, int64_t *launches_per_kind) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < kProfKinds; ++k) { ms_per_kind[k] = 0.0; launches_per_kind[k] = 0; }
  for (auto &e : g_prof) {
    float ms = 0.f;
    OCC_CUDA(cudaEventSynchronize(e.b));
    OCC_CUDA(cudaEventElapsedTime(&ms, e.a, e.b));
    ms_per_kind[e.kind] += ms;
    launches_per_kind[e.kind] += 1;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  g_prof.clear();
  return 0;
}

extern "C" int occb200_point_cloud_to_range_image_idx(const double *points, int B, int64_t N, const float *v2l,
                                                      const float *azc, const float *incl, int H, int W,
                                                      int64_t *ri_idx, double *ri_range, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(B >= 0 && N >= 0 && H >= 1 && W >= 1, "bad sizes");
  if (B == 0 || N == 0) return 0;
  dim3 grid((unsigned)ceil_div(N, 256), (unsigned)B);
  // mono = 0: the operator accepts arbitrary tables, as the reference does (linear argmin)
  k_project_points<<<grid, 256, 0, stream>>>(points, B, N, v2l, azc, incl, H, W, 0, ri_idx, ri_range);
  OCC_KERNEL_OK("k_project_points");
  return 0;
}

extern "C" void occb200_host_pose_pack(const float *boxes7, const float *trig4, int64_t n, occb200_pose_t *poses) {
  for (int64_t i = 0; i < n; ++i) {
    occb200_pose_t &p = poses[i];
    for (int k = 0; k < 7; ++k) p.box[k] = boxes7[7 * i + k];
    const float a = (float)((double)boxes7[7 * i + 6] + M_PI / 2);   // points_in_boxes_cpu.cpp:19
    p.cos_pib = cosf(a);
    p.sin_pib = sinf(a);
    p.cos_m = trig4[4 * i + 0];
    p.sin_m = trig4[4 * i + 1];
    p.cos_p = trig4[4 * i + 2];
    p.sin_p = trig4[4 * i + 3];
    p.pad[0] = p.pad[1] = p.pad[2] = 0.f;
  }
}

extern "C" void occb200_host_box_trig(const float *boxes7, int64_t n, float *trig) {
  for (int64_t i = 0; i < n; ++i) {
    const float a = (float)((double)boxes7[7 * i + 6] + M_PI / 2);
    trig[2 * i] = cosf(a);
    trig[2 * i + 1] = sinf(a);
  }
}
