// annotate.cu -- batched tracklet annotation: crop -> box frame -> voxelise -> visibility.
//
// Reference path: tools/occ/occ_annotate.py get_local_point_list :91-138 and
// OccAnnotator.annotate_trk :344-568 (normative step list: SURVEY.md Appendix A).
//
// Kernels (all on the caller's stream, no host synchronisation):
//   k_tracklet_presetup  one warp per tracklet: optimistic grid (box size = max over all frames)
//   k_frame_voxelize  one CTA per tracklet-frame: in-box -> box frame -> quantise -> bitset (A1/A2/A3)
//   k_tracklet_setup  one warp per tracklet: box size = max over KEPT frames; rare re-voxelisation (A2/A3)
//   k_scan_chunks     one CTA: exclusive scan of per-tracklet work chunks -> work list
//   k_table_setup     one CTA per (sensor frame, LiDAR): inclination row boundaries + lookup table
//   k_pair_build      one CTA per tracklet, one thread per (frame, LiDAR): voxel-index -> sensor-frame affine map,
//   k_visibility_fast persistent CTAs over 32-voxel chunks: the range-image "ray-cast" in f32 with
//                     rigorous error margins; tests whose outcome is not certain are queued    (A4/A5)
//   k_visibility_recheck  the queued tests, re-evaluated with the reference's exact f64 arithmetic
//   k_visibility_f64  (flags bit 0) every test in exact f64 -- the slow, margin-free formulation
//
// Why the fast kernel is still bit-exact.  Per test the reference takes three discrete decisions from
// f64 quantities: the range-image row (nearest inclination), the column (rounded azimuth) and
// `range_image >= range`.  The fast kernel evaluates the same quantities in f32 from a per-(frame,
// LiDAR) affine map of the voxel index, with an explicit bound on |f32 value - reference f64 value|
// (derived at each use below).  A decision is accepted only if it stays the same anywhere inside that
// bound; otherwise the test goes to a queue and k_visibility_recheck redoes it with project_exact().
// Labels therefore never depend on f32 rounding; only the amount of rechecked work does.
#include <math.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "geom.cuh"

namespace occb200 {

constexpr int kChunk = 256;         // f64 kernel: voxels per work item == threads per CTA
#ifndef OCC_VPL
#define OCC_VPL 2
#endif
#ifndef OCC_MINB
#define OCC_MINB 4
#endif
constexpr int kVPL = OCC_VPL;       // fast kernel, phase 2: voxels per lane
#ifndef OCC_VPL1
#define OCC_VPL1 2
#endif
constexpr int kVPL1 = OCC_VPL1;     // fast kernel, phase 1: voxels per lane
constexpr int kFastChunk = 32 * kVPL;   // fast kernel: voxels per work item
constexpr int kFastWarps = 8;       // fast kernel: independent warps per CTA
#ifndef OCC_PPI
#define OCC_PPI 16
#endif
constexpr int kPairsPerItem = OCC_PPI;   // fast kernel: (frame, LiDAR) pairs one work item covers for its 64 voxels
#ifndef OCC_P1
#define OCC_P1 8
#endif
constexpr int kPhase1Pairs = OCC_P1;     // pairs tested on EVERY non-occupied voxel before the survivors are compacted
constexpr float kAtanErr = 2.0e-6f;     // bound on |atan2_fast - atan2| (derivation at atan2_fast)
constexpr int kFrameStride = 8;          // pair order: frames 0,8,16,.. then 1,9,17,.. (spread viewpoints come first)
constexpr int kLutPerRow = 64;      // lookup-table cells reserved per inclination-table entry
constexpr int kTileR = 8, kTileC = 32;   // range-image max-pyramid tile (rows x columns)
#ifndef OCC_FT
#define OCC_FT 256
#endif
#ifndef OCC_FMINB
#define OCC_FMINB 6
#endif
constexpr int kFrameThreads = OCC_FT;
constexpr int kSmemBitWords = 8192; // 32 KB: grids up to 262 144 voxels keep their bitset in shared memory; the
                                    // launch asks only for what the largest grid of the batch needs

struct TrkGrid {
  int32_t dims[3];
  int32_t status;
  float mb[3];      // min bound of the canonical box (f32)
  int32_t flags;    // bit0: some point survived the q<dims filter, bit1: index error
  int64_t V;
  int64_t bits_off; // word offset of the occupancy bitset
  int32_t nchunks;
  int32_t B;
  int32_t redo;     // 1: the optimistic grid was wrong, k_frame_voxelize runs again for this tracklet
  int32_t pad;
};

// Per (sensor frame, LiDAR) constants of the fast path.  64 bytes, 4 x LDG.128.
// Row lookup happens in "u space": u(inc) = sin/(|sin| + cos), monotone with slope in [1/2, 1] over
// [-90, 90] degrees, so row boundaries never bunch up (unlike tan or sin).
struct __align__(16) SensCoef {
  float azc;
  float kcol;       // W / (2 pi)
  float Wf;
  float c_col;      // evaluation error of the f32 column coordinate (pixels)
  float u_lo, inv_w;   // lookup cell k covers [u_lo + k/inv_w, u_lo + (k+1)/inv_w); the cells span all of [-1, 1]
  int32_t ncell;
  int32_t H;
  int32_t W;
  int32_t ok;       // 0: no fast path for this sensor (table not strictly descending, H < 2, ...)
  int32_t tab_off;  // == incl_off: the table sits at 2*tab_off+1 in ub_pool (sentinels around it) and at
                    // kLutPerRow*tab_off in lut_pool
  float cell0;      // -u_lo * inv_w: cell = int(fma(u, inv_w, cell0))
  int64_t ri_off;
  float col0;       // W/2 - 0.5: colf = fma(az, -kcol, col0)
  float pad1;
};
static_assert(sizeof(SensCoef) == 64, "SensCoef must be 64 bytes");

// One (tracklet-frame, LiDAR) pair: p_sensor = A * (x,y,z voxel index) + b.  64 bytes, 4 x LDG.128.
struct __align__(16) PairCoef {
  float A[9];
  float b[3];
  float eps;        // bound on |f32 p - reference f64 p| per component (metres); < 0: no fast path
  int32_t sens;     // index into the SensCoef table
  int32_t q;        // pair index inside the tracklet: frame * L + LiDAR
  int32_t cull;     // 1: no voxel of the tracklet can be free through this pair (see make_pair)
};
static_assert(sizeof(PairCoef) == 64, "PairCoef must be 64 bytes");

// What one iteration of the fast visibility kernel needs of a surviving pair, in ONE 128-byte record (one L1 line,
// 8 x LDG.128, prefetched one iteration ahead): the pair's affine map, the fields of its sensor entry and the
// per-pair constants of the margins.  Built by k_pair_build.
struct __align__(16) PairHot {
  float A[9];
  float b[3];
  float eps;        // < 0: no fast path for this pair (every test goes to the exact recheck)
  int32_t q;
  float azc, inv_w, cell0m, col0;   // cell0m = cell0 - 0.5: the cell index is taken by round-to-nearest
  int32_t W, tab_off;
  int64_t ri_off;
  float e15;        // 1.5 eps
  float c1, c2;     // range margin: m = r * (r * 6e-7 + c1) + c2,  c1 = 2.01 sqrt(3) eps, c2 = 3 eps^2
  float ecol;       // (kAtanErr + 3e-7) * kcol + c_col
  float e15k;       // 1.5 eps * kcol
  float nkcol;      // -kcol
  uint32_t last;    // H - 1
  uint32_t ncm1;    // ncell - 1
  int32_t pad[2];
};
static_assert(sizeof(PairHot) == 128, "PairHot must be 128 bytes");

// What the fast visibility kernel needs of a tracklet, in one 64-byte record (4 x LDG.128).
struct __align__(16) TrkHot {
  int32_t V, dY, dZ;
  int32_t status;       // final status (flags of k_frame_voxelize folded in)
  int32_t nact;         // pairs that survived culling
  int32_t pad0;
  int64_t bits_off;
  int64_t label_off;
  int64_t pairs_base;   // first PairCoef of the tracklet in pairs_c
  int64_t pad1[2];
};
static_assert(sizeof(TrkHot) == 64, "TrkHot must be 64 bytes");

struct Workspace {
  TrkGrid *grids;        // [T]
  int32_t *frame_kept;   // [F]
  int32_t *frame_trk;    // [F]
  int32_t *redo_list;    // [F] frames of tracklets whose optimistic grid was wrong
  unsigned long long *redo_count;
  int64_t *chunk_off;    // [T+1]
  unsigned long long *pyr_flag;  // [0] 1: pyramid built (fits), culling enabled -- written on the side stream
  unsigned long long *counter;   // [0] phase-1 ticket, [1] recheck-queue length, [3] phase-1 items,
                                 // [4] phase-2 ticket, [5] phase-2 items
  uint32_t *bits;        // occupancy bitsets
  int64_t bits_words;
  SensCoef *sens;        // [SF*L]
  float *ub_pool;        // [2*incl_len+2] u-space row boundaries, table at 2*incl_off+1 with a sentinel on each side
  uint16_t *lut_pool;    // [incl_len * kLutPerRow]
  int32_t *tab_claim;    // [incl_len] 1 at a table's offset once a CTA has taken on building its lookup table
  int64_t incl_len;
  int4 *queue;           // recheck queue: (tracklet, voxel, pair index q = i*L + c, unused)
  int64_t queue_cap;
  PairHot *pairs_c;      // [F*L] the non-culled pairs of each tracklet, compacted at trk_frame_off[t] * L
  TrkHot *hot;           // [T]
  uint32_t *free_bits;   // same layout as `bits`: voxels proven free
  int2 *item_map;        // [items_cap] work items of the fast kernel: (tracklet, chunk | slice << 20)
  int64_t items_cap;
  int32_t *unk_list;     // [total] per tracklet (at label_off): voxels still undecided after phase 1, any order
  uint32_t *n_unk;       // [T] length of each tracklet's list
  int64_t *pyr_off;      // [SF*L + 1] first tile of each range image in pyr
  float *pyr;            // [pyr_tiles] max of the range image over tiles of kTileR x kTileC pixels
  int64_t pyr_tiles;
};

static int64_t ws_layout(int32_t T, int64_t F, int64_t total, int64_t SF, int32_t L, int64_t incl_len,
                         int64_t pyr_tiles, int64_t items_cap, char *base, Workspace *w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  int64_t o_grid = take(sizeof(TrkGrid) * (int64_t)T);
  int64_t o_kept = take(4 * F);
  int64_t o_ftrk = take(4 * F);
  int64_t o_redo = take(4 * F);
  int64_t o_rc = take(8);
  int64_t o_choff = take(8 * ((int64_t)T + 1));
  int64_t o_cnt = take(8 * 8);
  int64_t o_pf = take(8);
  int64_t words = total / 32 + T + 1;
  int64_t o_bits = take(4 * words);
  int64_t o_fbits = take(4 * words);
  int64_t o_imap = take(8 * items_cap);
  int64_t o_tab = take(sizeof(SensCoef) * SF * L);
  int64_t o_ub = take(4 * (2 * incl_len + 2));
  int64_t o_lut = take(2 * incl_len * kLutPerRow);
  int64_t o_claim = take(4 * std::max<int64_t>(incl_len, 1));
  // recheck queue: ~0.4 % of the EXECUTED tests (~0.1 % of the nominal ones) are undecided in f32; room for 1/64
  // of the nominal tests, bounded.  Tests beyond the capacity are decided in place (exact_from_ids).
  const double nominal = (double)total * (T > 0 ? (double)F / T : 0.0) * L;
  int64_t qcap = (int64_t)std::min(std::max(nominal / 64.0, 65536.0), 16.0 * 1024 * 1024);
  int64_t o_q = take(16 * qcap);
  int64_t o_pc = take(sizeof(PairHot) * F * L);
  int64_t o_na = take(sizeof(TrkHot) * (int64_t)T);
  int64_t o_po = take(8 * (SF * L + 1));
  int64_t o_py = take(4 * pyr_tiles);
  int64_t o_ul = take(4 * total);
  int64_t o_nu = take(4 * (int64_t)T);
  if (w) {
    w->unk_list = (int32_t *)(base + o_ul);
    w->n_unk = (uint32_t *)(base + o_nu);
    w->pairs_c = (PairHot *)(base + o_pc);
    w->hot = (TrkHot *)(base + o_na);
    w->pyr_off = (int64_t *)(base + o_po);
    w->pyr = (float *)(base + o_py);
    w->pyr_tiles = pyr_tiles;
    w->grids = (TrkGrid *)(base + o_grid);
    w->frame_kept = (int32_t *)(base + o_kept);
    w->frame_trk = (int32_t *)(base + o_ftrk);
    w->redo_list = (int32_t *)(base + o_redo);
    w->redo_count = (unsigned long long *)(base + o_rc);
    w->chunk_off = (int64_t *)(base + o_choff);
    w->counter = (unsigned long long *)(base + o_cnt);
    w->pyr_flag = (unsigned long long *)(base + o_pf);
    w->bits = (uint32_t *)(base + o_bits);
    w->bits_words = words;
    w->free_bits = (uint32_t *)(base + o_fbits);
    w->item_map = (int2 *)(base + o_imap);
    w->items_cap = items_cap;
    w->sens = (SensCoef *)(base + o_tab);
    w->ub_pool = (float *)(base + o_ub);
    w->lut_pool = (uint16_t *)(base + o_lut);
    w->tab_claim = (int32_t *)(base + o_claim);
    w->incl_len = incl_len;
    w->queue = (int4 *)(base + o_q);
    w->queue_cap = qcap;
  }
  return off;
}

// ---------------------------------------------------------------------------------------------
// The sensor-side setup (row tables, range-image pyramid) does not depend on the tracklets, so it runs on
// a side stream concurrently with the crop/voxelise chain and joins before k_pair_build (fork/join with
// events; capturable in a CUDA graph).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream g_side[64];
static std::mutex g_side_mu;

static int side_stream(SideStream **out) {
  int dev = 0;
  OCC_CUDA(cudaGetDevice(&dev));
  OCC_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  std::lock_guard<std::mutex> lk(g_side_mu);
  SideStream &s = g_side[dev];
  if (!s.stream) {
    OCC_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    OCC_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    OCC_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Optional per-kernel timing (bench.py's roofline): CUDA events recorded around each kernel of the
// pipeline on the caller's stream; durations are summed per kernel when the profile is read.
enum { kProfInbox = 0, kProfSetup, kProfScan, kProfVoxelize, kProfVisibility, kProfPairSetup, kProfRecheck, kProfPairCull, kProfKinds };
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
// Crop + voxelise.  The grid of a tracklet depends on its box size = max over the frames that have at
// least one in-box point (occ_annotate.py:111-112, 132-133), which is only known after every frame has
// been cropped.  Almost always every frame has such a point, so the pipeline is optimistic:
//   k_tracklet_presetup  grid from the max over ALL frames
//   k_frame_voxelize     ONE pass over the points: in-box test, box frame, quantise, set bits; records
//                        which frames had in-box points
//   k_tracklet_setup     recomputes the size from the kept frames; if it differs the tracklet's bits are
//                        cleared and a second k_frame_voxelize pass redoes just that tracklet
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_from_size(TrkGrid &g, const float sz[3], float vsf, int64_t cap, int chunk) {
  g.status = OCCB200_OK;
  for (int k = 0; k < 3; ++k) g.dims[k] = (int)ceilf(__fdiv_rn(sz[k], vsf));   // :414-416
  g.mb[0] = __fmul_rn(sz[0], -0.5f);           // min over corners of [0,0,0,w,l,h,0] (:422-423)
  g.mb[1] = __fmul_rn(sz[1], -0.5f);
  g.mb[2] = __fmul_rn(sz[2], 0.0f);
  g.V = (int64_t)g.dims[0] * g.dims[1] * g.dims[2];
  if (g.V > cap || g.V <= 0 || g.V >= (1ll << 31)) {
    g.status = -1;                             // caller's slot too small: reported, nothing written
    g.V = 0;
  }
  g.nchunks = (int)((g.V + chunk - 1) / chunk);
}

// One in-box point -> bit index in the tracklet's occupancy bitset, or -1; flags |= 1 kept, |= 2 index error.
__device__ __forceinline__ int64_t voxel_of_point(const BoxTest &bt, const occb200_pose_t &ps, const TrkGrid &g,
                                                  float vsf, float x, float y, float z, int &flags) {
  if (!pt_in_box(bt, x, y, z)) return -1;
  flags |= 4;                                   // the frame has an in-box point
  // local = (p + (-origin)) @ [[c,-s,0],[s,c,0],[0,0,1]]  (:117-122, lidar_box3d.py:165-184):
  // sgemm accumulates k = 0,1,2 as an FMA chain; the k=2 terms are exact no-ops.
  const float c = ps.cos_m, s = ps.sin_m;       // torch f32 cos/sin(-yaw)
  const float tx = __fadd_rn(x, -ps.box[0]), ty = __fadd_rn(y, -ps.box[1]), tz = __fadd_rn(z, -ps.box[2]);
  const float lx = __fmaf_rn(ty, s, __fmul_rn(tx, c));
  const float ly = __fmaf_rn(ty, c, __fmul_rn(tx, -s));
  const float lz = tz;
  // q = floor((local - min_bound) / vs)  (:425)
  float qx = floorf(__fdiv_rn(__fsub_rn(lx, g.mb[0]), vsf));
  float qy = floorf(__fdiv_rn(__fsub_rn(ly, g.mb[1]), vsf));
  float qz = floorf(__fdiv_rn(__fsub_rn(lz, g.mb[2]), vsf));
  const float dX = (float)g.dims[0], dY = (float)g.dims[1], dZ = (float)g.dims[2];
  if (!(qx < dX && qy < dY && qz < dZ)) return -1;   // only the upper bound is filtered (:430-431)
  flags |= 1;
  if (qx < 0.f) qx += dX;                       // PyTorch negative-index wrap (:436)
  if (qy < 0.f) qy += dY;
  if (qz < 0.f) qz += dZ;
  if (qx < 0.f || qy < 0.f || qz < 0.f) {
    flags |= 2;                                 // IndexError in the reference
    return -1;
  }
  return ((int64_t)qx * g.dims[1] + (int64_t)qy) * g.dims[2] + (int64_t)qz;
}

__global__ void __launch_bounds__(256)
k_tracklet_presetup(int T, const int64_t *__restrict__ trk_frame_off, const occb200_pose_t *__restrict__ poses,
                    const int64_t *__restrict__ frame_pt_off, const int64_t *__restrict__ label_off, float vsf, int chunk, TrkGrid *__restrict__ grids,
                    int32_t *__restrict__ frame_trk, unsigned long long *__restrict__ redo_count,
                    unsigned long long *__restrict__ counter, uint32_t *__restrict__ n_unk,
                    int64_t *__restrict__ n_unknown, int64_t *__restrict__ n_steps) {
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per tracklet
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *redo_count = 0ull;
  if (blockIdx.x == 0 && threadIdx.x < 8) counter[threadIdx.x] = 0ull;
  if (t >= T) return;
  const int64_t f0 = trk_frame_off[t], f1 = trk_frame_off[t + 1];
  float sz[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t f = f0 + lane; f < f1; f += 32) {
    frame_trk[f] = t;
    if (frame_pt_off[f + 1] > frame_pt_off[f])     // a frame without candidates cannot be a kept frame
#pragma unroll
      for (int k = 0; k < 3; ++k) sz[k] = fmaxf(sz[k], poses[f].box[3 + k]);
  }
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 3; ++k) sz[k] = fmaxf(sz[k], __shfl_xor_sync(0xffffffffu, sz[k], o));
  if (lane != 0) return;
  TrkGrid g;
  g.B = (int)(f1 - f0);
  g.flags = 0;
  g.nchunks = 0;
  g.V = 0;
  g.redo = 0;
  g.pad = 0;
  g.bits_off = label_off[t] / 32 + t;
  g.dims[0] = g.dims[1] = g.dims[2] = 0;
  g.mb[0] = g.mb[1] = g.mb[2] = 0.f;
  if (g.B < 10) g.status = OCCB200_SKIP_SHORT;  // :344
  else if (sz[0] == -INFINITY) g.status = OCCB200_NO_POINTS;   // no candidate point at all (:129)
  else grid_from_size(g, sz, vsf, label_off[t + 1] - label_off[t], chunk);
  grids[t] = g;
  n_unknown[t] = 0;
  n_unk[t] = 0u;
  if (n_steps) n_steps[t] = 0;
}

#ifndef OCC_PPT
#define OCC_PPT 4
#endif
constexpr int kPtsPerThread = OCC_PPT;          // independent point loads in flight per thread
__global__ void __launch_bounds__(kFrameThreads, OCC_FMINB)
k_frame_voxelize(const occb200_pose_t *__restrict__ poses, const float *__restrict__ points, int stride,
                 const int64_t *__restrict__ frame_pt_off, int32_t *__restrict__ frame_kept,
                 const int32_t *__restrict__ frame_trk, TrkGrid *__restrict__ grids,
                 uint32_t *__restrict__ bits, float vsf, const int32_t *__restrict__ redo_list,
                 const unsigned long long *__restrict__ redo_count, int smem_words) {
  extern __shared__ uint32_t s_bits[];            // smem_words words
  __shared__ int s_flags;
  // first pass: CTA b = tracklet-frame b.  second pass (redo_list != NULL): a small grid strides over the
  // frames of the tracklets whose optimistic grid was wrong -- usually none.
  const bool redo_pass = redo_list != nullptr;
  const long long n_work = redo_pass ? (long long)*redo_count : (long long)gridDim.x;
  for (long long wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
  __syncthreads();
  const int64_t f = redo_pass ? (int64_t)redo_list[wi] : (int64_t)wi;
  const int t = frame_trk[f];
  const TrkGrid g = grids[t];
  if (g.status != OCCB200_OK) {
    if (threadIdx.x == 0 && !redo_pass) frame_kept[f] = 0;
    continue;
  }
  const int words = (int)((g.V + 31) / 32);
  const bool use_smem = words <= smem_words;
  uint32_t *gbits = bits + g.bits_off;
  if (use_smem)
    for (int w = threadIdx.x; w < words; w += kFrameThreads) s_bits[w] = 0u;
  if (threadIdx.x == 0) s_flags = 0;
  __syncthreads();

  const occb200_pose_t ps = poses[f];
  const BoxTest bt = make_box_test(ps.box, ps.cos_pib, ps.sin_pib);
  const int64_t n0 = frame_pt_off[f], n1 = frame_pt_off[f + 1];
  const int lane = threadIdx.x & 31;
  int flags = 0;
  for (int64_t base = n0; base < n1; base += kFrameThreads * kPtsPerThread) {   // warp-uniform trip count
    float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
#pragma unroll
    for (int u = 0; u < kPtsPerThread; ++u) {
      const int64_t j = base + u * kFrameThreads + threadIdx.x;
      const float *p = points + j * stride;
      const bool ok = j < n1;
      px[u] = ok ? ld_stream(p) : 0.f;
      py[u] = ok ? ld_stream(p + 1) : 0.f;
      pz[u] = ok ? ld_stream(p + 2) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kPtsPerThread; ++u) {
      const int64_t j = base + u * kFrameThreads + threadIdx.x;
      int64_t idx = -1;
      if (j < n1) idx = voxel_of_point(bt, ps, g, vsf, px[u], py[u], pz[u], flags);
      int word = -1;
      uint32_t bit = 0u;
      if (idx >= 0) {
        word = (int)(idx >> 5);
        bit = 1u << (idx & 31);
        // most points land in voxels that are already marked: look before touching an atomic
        const uint32_t cur = use_smem ? s_bits[word] : __ldg(gbits + word);
        if (cur & bit) word = -1;
      }
      // warp-level dedup of what is left: lanes on the same bitset word merge their bits, one atomic per word
      if (__any_sync(0xffffffffu, word >= 0)) {
        const unsigned peers = __match_any_sync(0xffffffffu, word);
        const uint32_t merged = __reduce_or_sync(peers, word >= 0 ? bit : 0u);
        if (word >= 0 && lane == __ffs(peers) - 1) {
          if (use_smem) atomicOr(&s_bits[word], merged);
          else atomicOr(&gbits[word], merged);
        }
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
  if (threadIdx.x == 0) {
    const int fl = s_flags;
    if (!redo_pass) frame_kept[f] = (fl & 4) ? 1 : 0;
    if (fl & 3) atomicOr(&grids[t].flags, fl & 3);
  }
  }
}

// --save-mean-var support (occ_annotate.py:627-645): for every candidate point, its box-frame coordinates and raw
// quantised voxel coordinates with the tracklet's FINAL grid, exactly as k_frame_voxelize derives them.  Row
// (t, qx, qy, qz) for a point the reference keeps (in the box, q < dims), (-1, 0, 0, 0) otherwise.  Negative
// coordinates are NOT wrapped here: the reference groups by the raw values (sst_ops.py:150-181).
__global__ void __launch_bounds__(256)
k_frame_points(const occb200_pose_t *__restrict__ poses, const float *__restrict__ points, int stride,
               const int64_t *__restrict__ frame_pt_off, const int32_t *__restrict__ frame_trk,
               const TrkGrid *__restrict__ grids, const int32_t *__restrict__ status, float vsf,
               float *__restrict__ loc_out, int32_t *__restrict__ q_out) {
  const int64_t f = blockIdx.x;
  const int t = frame_trk[f];
  const TrkGrid g = grids[t];
  const bool live = status[t] == OCCB200_OK;
  const occb200_pose_t ps = poses[f];
  const BoxTest bt = make_box_test(ps.box, ps.cos_pib, ps.sin_pib);
  const float dX = (float)g.dims[0], dY = (float)g.dims[1], dZ = (float)g.dims[2];
  for (int64_t j = frame_pt_off[f] + threadIdx.x; j < frame_pt_off[f + 1]; j += blockDim.x) {
    const float *p = points + j * stride;
    const float x = p[0], y = p[1], z = p[2];
    float lx = 0.f, ly = 0.f, lz = 0.f;
    int4 row = make_int4(-1, 0, 0, 0);
    if (live && pt_in_box(bt, x, y, z)) {
      const float c = ps.cos_m, s = ps.sin_m;
      const float tx = __fadd_rn(x, -ps.box[0]), ty = __fadd_rn(y, -ps.box[1]), tz = __fadd_rn(z, -ps.box[2]);
      lx = __fmaf_rn(ty, s, __fmul_rn(tx, c));
      ly = __fmaf_rn(ty, c, __fmul_rn(tx, -s));
      lz = tz;
      const float qx = floorf(__fdiv_rn(__fsub_rn(lx, g.mb[0]), vsf));
      const float qy = floorf(__fdiv_rn(__fsub_rn(ly, g.mb[1]), vsf));
      const float qz = floorf(__fdiv_rn(__fsub_rn(lz, g.mb[2]), vsf));
      if (qx < dX && qy < dY && qz < dZ) row = make_int4(t, (int)qx, (int)qy, (int)qz);
    }
    loc_out[3 * j + 0] = lx;
    loc_out[3 * j + 1] = ly;
    loc_out[3 * j + 2] = lz;
    reinterpret_cast<int4 *>(q_out)[j] = row;
  }
}

__global__ void __launch_bounds__(256)
k_tracklet_setup(int T, const int64_t *__restrict__ trk_frame_off, const occb200_pose_t *__restrict__ poses,
                 const int64_t *__restrict__ frame_pt_off, const int32_t *__restrict__ frame_kept,
                 const int64_t *__restrict__ label_off, float vsf, int chunk,
                 TrkGrid *__restrict__ grids, uint32_t *__restrict__ bits, int32_t *__restrict__ redo_list,
                 unsigned long long *__restrict__ redo_count, int32_t *__restrict__ dims_out,
                 float *__restrict__ sizes_out, int32_t *__restrict__ status_out) {
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per tracklet
  const int lane = threadIdx.x & 31;
  if (t >= T) return;
  const int64_t f0 = trk_frame_off[t], f1 = trk_frame_off[t + 1];
  TrkGrid g = grids[t];
  float sz[3] = {-INFINITY, -INFINITY, -INFINITY}, sz_all[3] = {-INFINITY, -INFINITY, -INFINITY};
  int kept = 0;
  for (int64_t f = f0 + lane; f < f1; f += 32) {
    const bool k_ = frame_kept[f] != 0;           // occ_annotate.py:111-112, :132-133 (box_mode="max")
    kept += k_ ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float b = poses[f].box[3 + k];
      if (frame_pt_off[f + 1] > frame_pt_off[f]) sz_all[k] = fmaxf(sz_all[k], b);   // what presetup assumed
      if (k_) sz[k] = fmaxf(sz[k], b);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    kept += __shfl_xor_sync(0xffffffffu, kept, o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sz[k] = fmaxf(sz[k], __shfl_xor_sync(0xffffffffu, sz[k], o));
      sz_all[k] = fmaxf(sz_all[k], __shfl_xor_sync(0xffffffffu, sz_all[k], o));
    }
  }
  if (g.status == OCCB200_OK || g.status == -1) {
    if (kept == 0) {
      g.status = OCCB200_NO_POINTS;               // :129
      g.V = 0;
      g.nchunks = 0;
      g.dims[0] = g.dims[1] = g.dims[2] = 0;
    } else if (sz[0] != sz_all[0] || sz[1] != sz_all[1] || sz[2] != sz_all[2]) {
      // a frame without in-box points carried the largest box: clear the bits set with the optimistic grid
      // and let the second k_frame_voxelize pass redo this tracklet with the true one.
      const int old_words = (int)((g.V + 31) / 32);
      grid_from_size(g, sz, vsf, label_off[t + 1] - label_off[t], chunk);
      g.flags = 0;
      g.redo = 1;
      uint32_t *gbits = bits + g.bits_off;
      for (int w = lane; w < old_words; w += 32) gbits[w] = 0u;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(redo_count, (unsigned long long)kept);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int64_t f = f0 + lane; f - lane < f1; f += 32) {        // kept frames, warp-compacted
        const bool k_ = f < f1 && frame_kept[f] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k_);
        if (k_) redo_list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)f;
        base += __popc(m);
      }
    }
  }
  if (lane != 0) return;
  grids[t] = g;
  for (int k = 0; k < 3; ++k) {
    dims_out[3 * t + k] = g.dims[k];
    sizes_out[3 * t + k] = (g.status == OCCB200_OK) ? sz[k] : 0.f;
  }
  status_out[t] = g.status;                    // refined by the visibility kernel (flags)
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
  for (int t = a; t < b; ++t) sum += (grids[t].status == OCCB200_OK) ? grids[t].nchunks : 0;
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
    run += (grids[t].status == OCCB200_OK) ? grids[t].nchunks : 0;
  }
  if (tid == 1023) chunk_off[T] = s_part[1023];
  if (tid < 8) counter[tid] = 0ull;
}

// ---------------------------------------------------------------------------------------------
// Exact visibility test of one voxel centre against one (frame, LiDAR): occ_annotate.py:490-547.
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

// centre = coord.f64 * vs + min_bound + vs/2, left to right (:467-471)
__device__ __forceinline__ void voxel_centre(const TrkGrid &g, int64_t f, double vs, double &cx, double &cy,
                                             double &cz) {
  const int YZ = g.dims[1] * g.dims[2];
  const int x = (int)(f / YZ), y = (int)((f / g.dims[2]) % g.dims[1]), z = (int)(f % g.dims[2]);
  cx = __dadd_rn(__dadd_rn(__dmul_rn((double)x, vs), (double)g.mb[0]), vs / 2);
  cy = __dadd_rn(__dadd_rn(__dmul_rn((double)y, vs), (double)g.mb[1]), vs / 2);
  cz = __dadd_rn(__dadd_rn(__dmul_rn((double)z, vs), (double)g.mb[2]), vs / 2);
}

__device__ __noinline__ bool exact_test(double cx, double cy, double cz, const occb200_pose_t *__restrict__ ps,
                                        const occb200_sensor_t *__restrict__ sn, const float *__restrict__ incl_pool,
                                        const float *__restrict__ ri_pool) {
  const double rc = (double)__ldg(&ps->cos_p), rs = (double)__ldg(&ps->sin_p);   // :490-496
  // ego = centre @ [[c,-s,0],[s,c,0],[0,0,1]] + origin (:497-498); the z row is exact
  const double ex = __dadd_rn(__fma_rn(cy, rs, __dmul_rn(cx, rc)), (double)__ldg(&ps->box[0]));
  const double ey = __dadd_rn(__fma_rn(cy, rc, __dmul_rn(cx, -rs)), (double)__ldg(&ps->box[1]));
  const double ez = __dadd_rn(cz, (double)__ldg(&ps->box[2]));
  const SensorView sv = load_sensor(sn, incl_pool);
  int row, col;
  double rng;
  project_exact(ex, ey, ez, sv, row, col, rng);
  if (col < 0) col += sv.W;                      // negative index wraps (:543)
  const float ri = __ldg(ri_pool + __ldg(&sn->ri_off) + (int64_t)row * sv.W + col);
  return (double)ri >= rng;                      // :547
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
      double cx, cy, cz;
      voxel_centre(g, active ? f : 0, vs, cx, cy, cz);
      const int64_t f0 = trk_frame_off[t];
      for (int c = 0; c < L; ++c) {              // LiDARs (:525), OR-ed (:552-556)
        for (int i = 0; i < g.B; ++i) {          // frames (:479), OR-ed (:550)
          if (__all_sync(0xffffffffu, !need || is_free)) break;   // warp-uniform early exit
          if (need && !is_free) {
            const occb200_sensor_t *sn = sensors + (int64_t)__ldg(frame_sf + f0 + i) * L + c;
            is_free = exact_test(cx, cy, cz, poses + f0 + i, sn, incl_pool, ri_pool);
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

// ---------------------------------------------------------------------------------------------
// Fast path, part 1: per-table row lookup.  Table t_0 > t_1 > ... (flipped inclinations,
// occ_annotate.py:528).  argmin_h |inc - t_h| (first index on ties, :168-173) == number of
// midpoints m_h = (t_h + t_{h+1})/2 that lie above inc.  Boundaries are stored as ub_h = u(m_h).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double u_of_angle(double a) {
  const double s = sin(a), c = cos(a);
  return s / (fabs(s) + c);
}

__global__ void __launch_bounds__(256)
k_table_setup(int64_t n_sensors, const occb200_sensor_t *__restrict__ sensors, const float *__restrict__ incl_pool,
              SensCoef *__restrict__ sens, float *__restrict__ ub_pool, uint16_t *__restrict__ lut_pool,
              int32_t *__restrict__ tab_claim) {
  __shared__ float s_min[256];
  __shared__ SensCoef s_info;
  __shared__ int s_builder;
  const int64_t e = blockIdx.x;
  if (e >= n_sensors) return;
  const occb200_sensor_t &sn = sensors[e];
  const int H = sn.H;
  const int64_t off = sn.incl_off;
  const float *tab = incl_pool + off;
  float *ub = ub_pool + 2 * off + 1;              // ub[-1] = +2 and ub[H-1] = -2 are sentinels
  uint16_t *lut = lut_pool + off * kLutPerRow;
  const bool candidate = (sn.incl_mono == -1) && H >= 2 && H < 65535 && off < (1ll << 24);
  float local_min = INFINITY;
  if (candidate) {
    for (int h = threadIdx.x; h < H - 1; h += blockDim.x) {
      const double m = 0.5 * ((double)tab[h] + (double)tab[h + 1]);
      ub[h] = (float)u_of_angle(m);
    }
    if (threadIdx.x == 0) {
      ub[-1] = 2.f;
      ub[H - 1] = -2.f;
    }
  }
  __syncthreads();
  if (candidate) {
    for (int h = threadIdx.x; h < H - 2; h += blockDim.x) local_min = fminf(local_min, ub[h] - ub[h + 1]);
    // the table must stay inside (-90, 90) degrees for u() to be monotone
    for (int h = threadIdx.x; h < H; h += blockDim.x)
      if (!(fabsf(tab[h]) < 1.5707f)) local_min = -1.f;
  }
  s_min[threadIdx.x] = local_min;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) s_min[threadIdx.x] = fminf(s_min[threadIdx.x], s_min[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    SensCoef sc;
    sc.azc = sn.azc;
    sc.kcol = (float)((double)sn.W / 6.28318530717958647692);
    sc.Wf = (float)sn.W;
    sc.col0 = 0.5f * (float)sn.W - 0.5f;
    // colf = fma(az, -kcol, W/2 - 0.5) with |az| <= 2 pi (no wrap: the column is taken modulo W afterwards): one
    // rounding at magnitude <= 1.5 W, the rounding of kcol (<= 2^-24 * W) and of az's sum (in kAtanErr's slack);
    // the reference wraps with float32(2 pi), which moves its colf by W * 2.8e-8 relative to an exact wrap
    sc.c_col = 3.0f * (float)sn.W * 1.1920929e-07f + (float)sn.W * 4e-8f;
    sc.u_lo = 0.f; sc.inv_w = 0.f; sc.ncell = 0; sc.cell0 = 0.f;
    sc.H = H; sc.W = sn.W; sc.ok = 0;
    sc.tab_off = (int32_t)off; sc.ri_off = sn.ri_off; sc.pad1 = 0.f;
    float spacing = s_min[0];
    if (candidate && H == 2) spacing = 0.25f;
    if (candidate && spacing > 1e-6f && isfinite(spacing) && sn.W >= 2 && sn.W < (1 << 22)) {
      // cells half as wide as the closest pair of boundaries (so a cell holds at most one), covering [-1, 1]
      const float w = 0.5f * spacing;
      const int ncell = (int)ceilf(2.04f / w) + 1;
      if (ncell <= H * kLutPerRow) {
        sc.u_lo = -1.02f;
        sc.inv_w = 1.0f / w;
        sc.cell0 = 1.02f * sc.inv_w;
        sc.ncell = ncell;
        sc.ok = 1;
      }
    }
    s_info = sc;
    sens[e] = sc;
    // the frames of a segment share one inclination table per LiDAR (same incl_off): every entry writes the same
    // boundaries above, but only the first CTA to claim the table builds its lookup cells
    s_builder = sc.ok ? (atomicCAS(tab_claim + off, 0, 1) == 0) : 0;
  }
  __syncthreads();
  const SensCoef sc = s_info;
  if (!sc.ok || !s_builder) return;
  // lut[k] = number of boundaries above the (slightly raised) upper end of cell k, i.e. the row of a point at
  // the top of the cell.  It only has to be a good starting guess: the kernel accepts a row only after checking
  // the two boundaries around it.
  const float w = 1.0f / sc.inv_w;
  for (int k = threadIdx.x; k < sc.ncell; k += blockDim.x) {
    const float ue = sc.u_lo + (float)(k + 1) * w + 0.02f * w;
    int lo = 0, hi = H - 1;                       // count of ub_h > ue (ub descending)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ub[mid] > ue) lo = mid + 1; else hi = mid;
    }
    lut[k] = (uint16_t)lo;
  }
}

// ---------------------------------------------------------------------------------------------
// Range-image max pyramid: pyr[tile] = max of the image over kTileR x kTileC pixels.  Used only to
// PROVE that a (frame, LiDAR) pair cannot free any voxel of a tracklet (all returns in the window
// the tracklet projects to are nearer than its nearest voxel), so that pair is skipped entirely.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_pyr_scan(int64_t n, const occb200_sensor_t *__restrict__ sensors, int64_t cap, int64_t *__restrict__ pyr_off,
           unsigned long long *__restrict__ counter) {
  __shared__ int64_t s_part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (n + 1023) / 1024;
  const int64_t a = min((int64_t)tid * per, n), b = min(a + per, n);
  auto tiles = [&](int64_t e) {
    const int H = sensors[e].H, W = sensors[e].W;
    return (int64_t)((H + kTileR - 1) / kTileR) * ((W + kTileC - 1) / kTileC);
  };
  int64_t sum = 0;
  for (int64_t e = a; e < b; ++e) sum += tiles(e);
  s_part[tid] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    int64_t v = (tid >= d) ? s_part[tid - d] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int64_t run = s_part[tid] - sum;
  for (int64_t e = a; e < b; ++e) {
    pyr_off[e] = run;
    run += tiles(e);
  }
  if (tid == 1023) {
    pyr_off[n] = s_part[1023];
    counter[0] = (s_part[1023] <= cap) ? 1ull : 0ull;      // 1: pyramid fits, culling enabled
  }
}

constexpr int kPyrRowGroups = 8;
__global__ void __launch_bounds__(256)
k_pyr_build(const occb200_sensor_t *__restrict__ sensors, const float *__restrict__ ri_pool,
            const int64_t *__restrict__ pyr_off, const unsigned long long *__restrict__ counter,
            float *__restrict__ pyr) {
  if (counter[0] == 0ull) return;
  const int e = blockIdx.x;                        // sensor entry; blockIdx.y = row group
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const occb200_sensor_t &sn = sensors[e];
  const int H = sn.H, W = sn.W;
  const int ntr = (H + kTileR - 1) / kTileR, ntc = (W + kTileC - 1) / kTileC;
  const float *img = ri_pool + sn.ri_off;
  float *out = pyr + pyr_off[e];
  constexpr int kU = 4;                            // tiles per warp in flight: 32 independent loads per lane
  // the image's tiles as one list, dealt to (row group, warp) in runs of kU: every warp is busy whatever the
  // image shape (a 200 x 600 image has only 19 tiles per tile row)
  const int ntile = ntr * ntc;
  if (((W & 1) == 0) && ((sn.ri_off & 1) == 0)) {
    // even width and offset: every pixel pair (col, col+1), col even, is 8-byte aligned.  A warp covers 64
    // columns = two tiles per row with one LDG.64 per lane: half the load instructions of the scalar path.
    const int npc = (ntc + 1) / 2;                 // tile pairs per tile row
    const int npair = ntr * npc;
    constexpr int kP = 2;                          // tile pairs per warp in flight: 16 LDG.64 per lane
    for (int i0 = (blockIdx.y * 8 + warp) * kP; i0 < npair; i0 += kPyrRowGroups * 8 * kP) {
      float2 v[kP][kTileR];
#pragma unroll
      for (int u = 0; u < kP; ++u) {
        const int p = i0 + u;
        const int tr = p / npc, tp = p - tr * npc;
        const int col = tp * (2 * kTileC) + 2 * lane;
#pragma unroll
        for (int r = 0; r < kTileR; ++r) {
          const int row = tr * kTileR + r;
          v[u][r] = (p < npair && col < W && row < H)
                        ? ld_stream2(reinterpret_cast<const float2 *>(img + (int64_t)row * W + col))
                        : make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kP; ++u) {
        const int p = i0 + u;
        const int tr = p / npc, tp = p - tr * npc;
        float m = 0.f;
#pragma unroll
        for (int r = 0; r < kTileR; ++r) m = fmaxf(m, fmaxf(v[u][r].x, v[u][r].y));
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));   // within 16 lanes
        const int tc = 2 * tp + (lane >> 4);       // lanes 0-15: first tile of the pair, 16-31: second
        if ((lane & 15) == 0 && p < npair && tc < ntc) out[tr * ntc + tc] = m;
      }
    }
    return;
  }
  for (int i0 = (blockIdx.y * 8 + warp) * kU; i0 < ntile; i0 += kPyrRowGroups * 8 * kU) {
    float v[kU][kTileR];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int tile = i0 + u;
      const int tr = tile / ntc, tc = tile - tr * ntc;
      const int col = tc * kTileC + lane;
#pragma unroll
      for (int r = 0; r < kTileR; ++r) {           // range images are >= 0 (0 = no return)
        const int row = tr * kTileR + r;
        v[u][r] = (tile < ntile && col < W && row < H) ? ld_stream(img + (int64_t)row * W + col) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float m = 0.f;
#pragma unroll
      for (int r = 0; r < kTileR; ++r) m = fmaxf(m, v[u][r]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0 && i0 + u < ntile) out[i0 + u] = m;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path, part 2: per (tracklet-frame, LiDAR) affine map, composed in f64 and stored in f32.
//   centre = vs*idx + c0,  c0 = min_bound + vs/2                    (:467-471)
//   ego    = Rm centre + o, Rm = [[c, s, 0], [-s, c, 0], [0, 0, 1]] (:490-498)
//   p      = V ego + tv                                              (:161-164)
//   =>  p = A idx + b,  A = vs V Rm,  b = V Rm c0 + V o + tv
// eps bounds |f32 chain - reference f64 chain| per component: the three FMAs round at most
// 3 * 2^-24 * M, the f32 coefficients contribute at most 2^-24 * M, M = |b| + sum |A| * max idx; the
// reference's own f64 roundings (~1e-14 m) vanish in the slack of the factor 6.
//
// Culling.  All voxel centres lie in the ball (centre pc = p(grid centre), radius R = half the grid
// diagonal).  Seen from the sensor the ball spans inclinations inc_c +- asin(R/d) and azimuths
// az_c +- asin(R/rho_c); the reference's row/column rules are monotone in those angles, so every
// pixel any centre can map to lies in the row/column window of the interval ends (padded by one).
// If the largest return in that window (from the tile pyramid) is below d - R, `ri >= range` is
// false for every voxel of the tracklet through this pair: the pair is dropped from the work list.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ PairCoef
make_pair(int64_t f, int c, int q, int L, const TrkGrid &g, const occb200_pose_t *__restrict__ poses,
          const int32_t *__restrict__ frame_sf, const occb200_sensor_t *__restrict__ sensors,
          const SensCoef *__restrict__ sens, double vs, const int64_t *__restrict__ pyr_off,
          const float *__restrict__ pyr, const uint16_t *__restrict__ lut_pool,
          const unsigned long long *__restrict__ counter) {
  const occb200_pose_t &ps = poses[f];
  const int64_t se = (int64_t)frame_sf[f] * L + c;
  const occb200_sensor_t &sn = sensors[se];
  PairCoef pc;
  const double rc = (double)ps.cos_p, rs = (double)ps.sin_p;
  const double Rm[9] = {rc, rs, 0, -rs, rc, 0, 0, 0, 1};
  double V[12];
  for (int k = 0; k < 12; ++k) V[k] = (double)sn.v2l[k];
  double VR[9];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) VR[3 * r + k] = V[4 * r] * Rm[k] + V[4 * r + 1] * Rm[3 + k] + V[4 * r + 2] * Rm[6 + k];
  const double c0[3] = {(double)g.mb[0] + vs / 2, (double)g.mb[1] + vs / 2, (double)g.mb[2] + vs / 2};
  const double o[3] = {(double)ps.box[0], (double)ps.box[1], (double)ps.box[2]};
  float eps = 0.f;
  double pcen[3], R2 = 0.0;
  for (int r = 0; r < 3; ++r) {
    const double b = VR[3 * r] * c0[0] + VR[3 * r + 1] * c0[1] + VR[3 * r + 2] * c0[2] + V[4 * r] * o[0] +
                     V[4 * r + 1] * o[1] + V[4 * r + 2] * o[2] + V[4 * r + 3];
    double M = fabs(b);
    pcen[r] = b;
    for (int k = 0; k < 3; ++k) {
      const double a = vs * VR[3 * r + k];
      const double span = (double)(g.dims[k] > 0 ? g.dims[k] - 1 : 0);
      pc.A[3 * r + k] = (float)a;
      M += fabs(a) * span;
      pcen[r] += 0.5 * span * a;
    }
    pc.b[r] = (float)b;
    eps = fmaxf(eps, (float)(6.0 * 5.9604644775390625e-08 * M));
  }
  for (int k = 0; k < 3; ++k) {                    // half diagonal of the centre grid in the sensor frame
    double col2 = 0.0;
    for (int r = 0; r < 3; ++r) col2 += VR[3 * r + k] * VR[3 * r + k];
    const double span = vs * (double)(g.dims[k] > 0 ? g.dims[k] - 1 : 0);
    R2 += col2 * span * span;
  }
  const bool ok = sens[se].ok && isfinite(eps) && se < (1ll << 31);
  pc.eps = ok ? eps : -1.f;
  pc.sens = (int32_t)se;
  pc.q = q;
  pc.cull = 0;
  // ---- cull test (conservative; any doubt keeps the pair)
  if (counter[0] != 0ull && g.status == OCCB200_OK && sn.incl_mono == -1 && sn.H >= 1 && sn.W >= 1) {
    // f32 geometry: the window is padded by 1e-4 rad (>> f32 error, << a pixel) and one extra row / column
    // the columns of VR are orthogonal only up to the f32 inverse: 1.001 covers it, +1 mm absolute
    const float R = 0.5f * sqrtf((float)R2) * 1.001f + 1e-3f;
    const float cxs = (float)pcen[0], cys = (float)pcen[1], czs = (float)pcen[2];
    const float rho = sqrtf(cxs * cxs + cys * cys);
    const float d = sqrtf(rho * rho + czs * czs);
    if (d > 1.25f * R) {
      const float rmin = d - R - 1e-3f * (1.f + d * 1e-3f);
      const float delta = asinf(R / d) + 1e-4f;
      const float inc_c = atan2f(czs, rho);
      // rows of the interval ends: the table lookup gives the row at the top of a cell (the true row is that
      // or the next one), which is all a conservative window needs; sensors without a table use every row
      int r0 = 0, r1 = sn.H - 1;
      const SensCoef sc = sens[se];
      if (sc.ok) {
        const uint16_t *lut = lut_pool + (int64_t)sc.tab_off * kLutPerRow;
        float sh, ch, sl, cl;
        sincosf(fminf(inc_c + delta, 1.5707f), &sh, &ch);
        sincosf(fmaxf(inc_c - delta, -1.5707f), &sl, &cl);
        const float u_hi = sh / (fabsf(sh) + ch) + 1e-5f;
        const float u_lo = sl / (fabsf(sl) + cl) - 1e-5f;
        const int c_hi = max(0, min((int)floorf(fmaf(u_hi, sc.inv_w, sc.cell0)) + 2, sc.ncell - 1));
        const int c_lo = max(0, min((int)floorf(fmaf(u_lo, sc.inv_w, sc.cell0)) - 2, sc.ncell - 1));
        r0 = max((int)lut[c_hi] - 1, 0);
        r1 = min((int)lut[c_lo] + 2, sn.H - 1);
      }
      const int W = sn.W;
      const int ntc = (W + kTileC - 1) / kTileC;
      long long c_lo = 0, c_hi = W - 1;            // column window, possibly beyond [0, W): taken modulo W
      if (rho > 1.05f * R) {
        const float daz = asinf(R / rho) + 1e-4f;
        const float az_c = atan2f(cys, cxs) + sn.azc;
        const float kc = (float)W / 6.2831853f;
        const float cf_lo = ((float)W - 0.5f) - (az_c + daz + 3.14159265f) * kc;
        const float cf_hi = ((float)W - 0.5f) - (az_c - daz + 3.14159265f) * kc;
        if (cf_hi - cf_lo + 6.0f < (float)W) {
          c_lo = (long long)floorf(cf_lo) - 2;
          c_hi = (long long)ceilf(cf_hi) + 2;
        }
      }
      // tile columns covering [c_lo, c_hi] modulo W: segment 1 = tiles [ta, tb], segment 2 (wrapped) = [0, tw]
      const float *pimg = pyr + pyr_off[se];
      const int tr0 = r0 / kTileR, tr1 = r1 / kTileR;
      int ta = 0, tb = ntc - 1, tw = -1;
      if (c_hi - c_lo + 1 < W) {
        const long long a0 = ((c_lo % W) + W) % W;       // first column, in [0, W)
        const long long len = c_hi - c_lo + 1;
        ta = (int)(a0 / kTileC);
        tb = (int)((min(a0 + len, (long long)W) - 1) / kTileC);
        if (a0 + len > W) tw = (int)((a0 + len - W - 1) / kTileC);
      }
      // four independent loads in flight per step (a serial max chain would pay one L2 latency per tile); the
      // scan stops as soon as one tile reaches rmin -- the pair is kept then
      float m = 0.f;
      for (int tr = tr0; tr <= tr1 && m < rmin; ++tr) {
        const float *prow = pimg + (int64_t)tr * ntc;
        for (int seg = 0; seg < 2; ++seg) {
          const int s0 = seg ? 0 : ta, s1 = seg ? tw : tb;
          for (int tc = s0; tc <= s1 && m < rmin; tc += 4) {
            const float v0 = prow[tc], v1 = prow[min(tc + 1, s1)], v2 = prow[min(tc + 2, s1)],
                        v3 = prow[min(tc + 3, s1)];
            m = fmaxf(fmaxf(m, fmaxf(v0, v1)), fmaxf(v2, v3));
          }
        }
      }
      if (m < rmin) pc.cull = 1;
    }
  }
  return pc;
}

// j-th frame of a tracklet of B frames in the order 0, S, 2S, .., 1, S+1, .. (S = kFrameStride)
__device__ __forceinline__ int strided_frame(int j, int B) {
#pragma unroll
  for (int r = 0; r < kFrameStride; ++r) {
    const int cnt = (B - r + kFrameStride - 1) / kFrameStride;
    if (j < cnt) return r + j * kFrameStride;
    j -= cnt;
  }
  return B - 1;                                   // not reached for j < B
}

// One CTA per tracklet, one thread per (frame, LiDAR) pair: the pair's affine map and cull decision (make_pair),
// then the surviving pairs are written (merged with their sensor entry into 128-byte records) to
// the front of the tracklet's slot -- frames in strided order, so that the first kPhase1Pairs pairs look at the
// object from spread-out viewpoints; thread 0 fixes the final status and writes the tracklet's hot record; all
// threads then emit the phase-1 work items (one per chunk of 32 * kVPL1 voxels).  Item ids come from an atomic counter, so their order across tracklets is arbitrary.
__global__ void __launch_bounds__(256)
k_pair_build(int T, int L, const int64_t *__restrict__ trk_frame_off, const int64_t *__restrict__ label_off,
             const TrkGrid *__restrict__ grids, const occb200_pose_t *__restrict__ poses,
             const int32_t *__restrict__ frame_sf, const occb200_sensor_t *__restrict__ sensors,
             const SensCoef *__restrict__ sens, double vs, const int64_t *__restrict__ pyr_off,
             const float *__restrict__ pyr, const uint16_t *__restrict__ lut_pool,
             const unsigned long long *__restrict__ pyr_flag, PairHot *__restrict__ pairs_c,
             TrkHot *__restrict__ hot, int2 *__restrict__ item_map, long long items_cap,
             unsigned long long *__restrict__ counter, int32_t *__restrict__ status_out) {
  __shared__ long long s_i0, s_nitems;
  __shared__ int s_cnt[8];
  const int t = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const TrkGrid g = grids[t];
  // the flags of a tracklet whose grid was right the first time are final; those of a corrected one are still
  // being rewritten by the redo pass (side stream) -- it is treated as OK here and k_labels folds its flags in
  int status = g.status;
  if (status == OCCB200_OK && !g.redo) {
    if (g.flags & 2) status = OCCB200_INDEX_ERROR;
    else if (!(g.flags & 1)) status = OCCB200_EMPTY_AFTER_FILTER;
  }
  const int64_t base = trk_frame_off[t] * L;
  const int B = (int)(trk_frame_off[t + 1] - trk_frame_off[t]);
  const int n = (status == OCCB200_OK) ? B * L : 0;
  int count = 0;                                    // surviving pairs so far (same value in every thread)
  for (int j0 = 0; j0 < n; j0 += 256) {             // 256 pairs per pass, one per thread
    const int j = j0 + threadIdx.x;
    int q = 0;
    if (j < n) {
      const int pf = j / L;
      q = strided_frame(pf, B) * L + (j - pf * L);
    }
    PairCoef pc;
    bool keep = false;
    if (j < n) {
      const int i = q / L;
      pc = make_pair(trk_frame_off[t] + i, q - i * L, q, L, g, poses, frame_sf, sensors, sens, vs, pyr_off, pyr,
                     lut_pool, pyr_flag);
      keep = pc.cull == 0;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_cnt[warp] = __popc(mask);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      before += (w < warp) ? s_cnt[w] : 0;
      total += s_cnt[w];
    }
    if (keep) {
      const int dst = count + before + __popc(mask & ((1u << lane) - 1u));
      const SensCoef sc = sens[pc.sens];
      PairHot p;
#pragma unroll
      for (int k = 0; k < 9; ++k) p.A[k] = pc.A[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) p.b[k] = pc.b[k];
      p.eps = pc.eps; p.q = pc.q;
      p.azc = sc.azc; p.inv_w = sc.inv_w; p.cell0m = sc.cell0 - 0.5f; p.col0 = sc.col0;
      p.W = sc.W; p.tab_off = sc.tab_off; p.ri_off = sc.ri_off;
      p.e15 = 1.5f * pc.eps;
      p.c1 = 2.01f * 1.7321f * pc.eps;
      p.c2 = 3.0003f * pc.eps * pc.eps;
      p.ecol = fmaf(kAtanErr + 3.0e-7f, sc.kcol, sc.c_col);
      p.e15k = p.e15 * sc.kcol;
      p.nkcol = -sc.kcol;
      p.last = (uint32_t)(sc.H - 1);
      p.ncm1 = (uint32_t)(sc.ncell - 1);
      p.pad[0] = p.pad[1] = 0;
      const float4 *src = reinterpret_cast<const float4 *>(&p);
      float4 *d4 = reinterpret_cast<float4 *>(pairs_c + base + dst);
#pragma unroll
      for (int k = 0; k < 8; ++k) d4[k] = src[k];
    }
    count += total;
    __syncthreads();                                // s_cnt is reused by the next pass
  }
  if (threadIdx.x == 0) {
    const long long nitems =
        (status == OCCB200_OK && count > 0) ? (long long)((g.V + 32 * kVPL1 - 1) / (32 * kVPL1)) : 0;
    s_nitems = nitems;
    s_i0 = nitems ? (long long)atomicAdd(counter + 3, (unsigned long long)nitems) : 0;
    TrkHot h;
    h.V = (int32_t)g.V; h.dY = g.dims[1]; h.dZ = g.dims[2];
    h.status = status; h.nact = count; h.pad0 = 0;
    h.bits_off = g.bits_off; h.label_off = label_off[t]; h.pairs_base = base;
    h.pad1[0] = h.pad1[1] = 0;
    hot[t] = h;
    status_out[t] = status;
  }
  __syncthreads();
  const long long i0 = s_i0, nitems = s_nitems;
  for (long long i = threadIdx.x; i < nitems; i += blockDim.x)
    if (i0 + i < items_cap) item_map[i0 + i] = make_int2(t, (int)i);
}

// After phase 1: one CTA per tracklet emits the phase-2 work items -- (64 listed voxels) x (slice of kPairsPerItem
// of the pairs phase 1 did not cover).  item_map is reused: phase 1 has finished with it.
__global__ void __launch_bounds__(256)
k_phase2_emit(const TrkHot *__restrict__ hot, const uint32_t *__restrict__ n_unk, int2 *__restrict__ item_map,
              long long items_cap, unsigned long long *__restrict__ counter) {
  __shared__ long long s_i0, s_nitems;
  __shared__ int s_nslice;
  const int t = blockIdx.x;
  if (threadIdx.x == 0) {
    const int nact = hot[t].nact;
    const long long nchunk = (hot[t].status == OCCB200_OK) ? ((long long)n_unk[t] + kFastChunk - 1) / kFastChunk : 0;
    const int nslice = nact > kPhase1Pairs ? (nact - kPhase1Pairs + kPairsPerItem - 1) / kPairsPerItem : 0;
    const long long nitems = nchunk * nslice;
    s_nslice = nslice;
    s_nitems = nitems;
    s_i0 = nitems ? (long long)atomicAdd(counter + 5, (unsigned long long)nitems) : 0;
  }
  __syncthreads();
  const long long i0 = s_i0, nitems = s_nitems;
  const int nslice = s_nslice;
  for (long long i = threadIdx.x; i < nitems; i += blockDim.x)
    if (i0 + i < items_cap) item_map[i0 + i] = make_int2(t, (int)(i / nslice) | ((int)(i % nslice) << 20));
}

// Approximate f32 primitives (flush-to-zero MUFU forms, <= 2 ulp): their error is part of every margin.
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// atan2 in f32: a * P(a^2), a = min/max in [0,1], P of degree 6 (|atan(a) - a P(a^2)| <= 2.5e-7 in exact
// arithmetic, checked on 2e6 points), plus Horner rounding (<= 4e-7), the approximate division
// (2 ulp of a: <= 2.4e-7) and the quadrant fix-ups (2 roundings at <= pi: 2.4e-7 each).
// Total < 1.4e-6 rad; kAtanErr = 2e-6 is the bound used for every margin below (occb200_selftest_atan2
// measures the actual maximum on the device).
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = mn * rcp_approx(mx);
  const float s = a * a;
  float p = 0.006811790633946657f;
  p = fmaf(p, s, -0.0336042158305645f);
  p = fmaf(p, s, 0.07962366938591003f);
  p = fmaf(p, s, -0.1323334276676178f);
  p = fmaf(p, s, 0.19807815551757812f);
  p = fmaf(p, s, -0.3331736922264099f);
  p = fmaf(p, s, 0.9999961256980896f);
  float r = p * a;
  if (ay > ax) r = 1.57079632679489661923f - r;
  if (x < 0.f) r = 3.14159265358979323846f - r;
  return (y < 0.f) ? -r : r;
}

// float -> nearest integer (ties to even) without the conversion unit: valid for |x| < 2^22; anything else
// (NaN included) gives an arbitrary integer, which every caller clamps before using it as an index.
constexpr float kMagic = 12582912.f;              // 1.5 * 2^23
__device__ __forceinline__ int magic_int(float biased) { return __float_as_int(biased) - 0x4B400000; }

// One fast test.  Returns 2 = certainly free, 0 = certainly not free, 1 = undecided (recheck in f64).
// `ub` points at the table's boundary 0 (sentinels at ub[-1] = +2 and ub[H-1] = -2).
__device__ __forceinline__ int fast_test(const PairHot &p, float x, float y, float z, const float *__restrict__ ub,
                                         const uint16_t *__restrict__ lut, const float *__restrict__ ri_img) {
  const float px = fmaf(z, p.A[2], fmaf(y, p.A[1], fmaf(x, p.A[0], p.b[0])));
  const float py = fmaf(z, p.A[5], fmaf(y, p.A[4], fmaf(x, p.A[3], p.b[1])));
  const float pz = fmaf(z, p.A[8], fmaf(y, p.A[7], fmaf(x, p.A[6], p.b[2])));
  const float s2 = fmaf(py, py, px * px);
  const float r2 = fmaf(pz, pz, s2);
  const float inv_rho = rsqrt_approx(s2);
  const float inv_r = rsqrt_approx(r2);

  // ---- row: u = pz / (|pz| + rho);  |u - u_ref| <= 1.42 eps / r  +  evaluation (~6 ulp of 1)
  const float u = pz * rcp_approx(fmaf(s2, inv_rho, fabsf(pz)));
  // the lookup cell is only a starting guess (checked against the boundaries below): nearest instead of floor
  const unsigned cell = min((unsigned)magic_int(fmaf(u, p.inv_w, p.cell0m) + kMagic), p.ncm1);
  const unsigned row0 = min((unsigned)__ldg(lut + cell), p.last);                  // row at the top of the cell
  const float *ubr = ub + row0;
  const float b_up = __ldg(ubr - 1), b_here = __ldg(ubr), b_dn = __ldg(ubr + 1);   // ub[H] is never selected
  const bool step = b_here > u;                               // a cell holds at most one boundary
  const unsigned row = row0 + (step ? 1u : 0u);
  const float below = step ? b_dn : b_here;                   // boundary between row and row + 1
  const float above = step ? b_here : b_up;                   // boundary between row - 1 and row
  // accepted only if u lies strictly between the two boundaries of `row`, by more than its error
  const bool ok_row = fminf(u - below, above - u) > fmaf(p.e15, inv_r, 1.5e-6f) && row <= p.last;

  // ---- column: az = atan2(py, px) + azc (not wrapped: |az| <= 2 pi and the column is taken modulo W);
  //      colf = (W - 0.5) - (az + pi) / (2 pi) * W  (:176-191);  |az - az_ref| <= kAtanErr + 1.42 eps / rho
  const float az = atan2_fast(py, px) + p.azc;
  const float colf = fmaf(az, p.nkcol, p.col0);
  const float cb = colf + kMagic;                             // |colf| <= 1.5 W < 2^22
  const float cr = cb - kMagic;                               // == rintf(colf)
  const bool ok_col = fabsf(colf - cr) + fmaf(p.e15k, inv_rho, p.ecol) < 0.5f;
  int col = magic_int(cb);
  col += (col < 0) ? p.W : 0;                                 // fmod(round(colf), W) (:191) and
  col -= (col >= p.W) ? p.W : 0;                              // negative index wrap (:543)
  col = min((unsigned)col, (unsigned)(p.W - 1));

  // ---- range: free iff ri >= |p_ref|;  | |p| - |p_ref| | <= sqrt(3) eps
  const float ri = __ldg(ri_img + (min(row, p.last) * (unsigned)p.W + (unsigned)col));
  const float r = r2 * inv_r;
  const float m = fmaf(r, fmaf(r, 6.0e-7f, p.c1), p.c2);      // margin on squared ranges
  const float ri2 = ri * ri;
  const bool yes = ri2 >= r2 + m, no = ri2 <= r2 - m;
  return (ok_row && ok_col && (yes || no)) ? (yes ? 2 : 0) : 1;
}

template <typename T16>
__device__ __forceinline__ T16 load128(const T16 *p) {         // 128-byte record, warp-uniform address
  T16 v;
  const float4 *src = reinterpret_cast<const float4 *>(p);
  float4 *dst = reinterpret_cast<float4 *>(&v);
#pragma unroll
  for (int k = 0; k < 8; ++k) dst[k] = __ldg(src + k);
  return v;
}

template <typename T16>
__device__ __forceinline__ T16 load64(const T16 *p) {          // 64-byte record, warp-uniform address
  T16 v;
  const float4 *src = reinterpret_cast<const float4 *>(p);
  float4 *dst = reinterpret_cast<float4 *>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = __ldg(src + k);
  return v;
}

// Rare path of the fast kernel (recheck queue full): decide one test exactly, from global memory only.
__device__ __noinline__ bool exact_from_ids(int t, int f, int q, int L, double vs, const TrkGrid *__restrict__ grids,
                                            const int64_t *__restrict__ trk_frame_off,
                                            const occb200_pose_t *__restrict__ poses,
                                            const int32_t *__restrict__ frame_sf,
                                            const occb200_sensor_t *__restrict__ sensors,
                                            const float *__restrict__ incl_pool, const float *__restrict__ ri_pool) {
  const TrkGrid g = grids[t];
  double cx, cy, cz;
  voxel_centre(g, f, vs, cx, cy, cz);
  const int i = q / L, c = q - i * L;
  const int64_t f0 = trk_frame_off[t];
  const occb200_sensor_t *sn = sensors + (int64_t)__ldg(frame_sf + f0 + i) * L + c;
  return exact_test(cx, cy, cz, poses + f0 + i, sn, incl_pool, ri_pool);
}

// Every warp is on its own: it claims a work item from an atomic counter, tests, and ORs the voxels it proved
// free into the global free bitset.  No shared memory, no barriers.  Two launches:
//   PHASE 1  item = 64 consecutive voxels of a tracklet x its first kPhase1Pairs surviving pairs.  Most voxels that
//            can be freed at all are freed here; what is left (non-occupied, not free) is appended to the
//            tracklet's list (warp-aggregated atomics, arbitrary order).
//   PHASE 2  item = 64 LISTED voxels x a slice of kPairsPerItem of the remaining pairs: every lane holds a voxel
//            that really needs the tests (dense lanes), and voxels freed in phase 1 are never tested again.
// Labels are written afterwards by k_labels from the occupancy and free bitsets.
template <int PHASE, int VPL>
__global__ void __launch_bounds__(32 * kFastWarps, OCC_MINB)
k_visibility_fast(int L, const int64_t *__restrict__ trk_frame_off,
                  const occb200_pose_t *__restrict__ poses, const int32_t *__restrict__ frame_sf,
                  const occb200_sensor_t *__restrict__ sensors, const float *__restrict__ incl_pool,
                  const float *__restrict__ ri_pool, double vs, const TrkGrid *__restrict__ grids,
                  unsigned long long *__restrict__ counter, long long items_cap,
                  const uint32_t *__restrict__ bits, uint32_t *__restrict__ free_bits,
                  const int2 *__restrict__ item_map, const TrkHot *__restrict__ hot,
                  const PairHot *__restrict__ pairs,
                  const float *__restrict__ ub_pool, const uint16_t *__restrict__ lut_pool,
                  int4 *__restrict__ queue, long long queue_cap, int32_t *__restrict__ unk_list,
                  uint32_t *__restrict__ n_unk, int64_t *__restrict__ n_steps) {
  const int lane = threadIdx.x & 31;
  unsigned long long *ticket = counter + (PHASE == 1 ? 0 : 4);
  const long long total = min((long long)counter[PHASE == 1 ? 3 : 5], items_cap);
  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(ticket, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const int2 m = __ldg(item_map + item);
    const int t = m.x, chunk = m.y & 0xfffff, slice = m.y >> 20;
    const TrkHot h = load64(hot + t);
    const int dZ = h.dZ, yz = h.dY * dZ;
    const int fbase = chunk * (32 * VPL);
    int32_t *list = unk_list + h.label_off;
    int vf[VPL];              // flat voxel index of this lane's voxel v
    float vx[VPL], vy[VPL], vz[VPL];
    unsigned todo = 0u;        // bit v: this lane's voxel v exists, holds no point and is not known to be free
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      bool active;
      int f;
      if (PHASE == 1) {
        f = fbase + 32 * v + lane;
        active = f < h.V;
        unsigned w = 0xffffffffu;
        const int64_t word = h.bits_off + (int64_t)chunk * VPL + v;
        if (fbase + 32 * v < h.V) w = __ldg(bits + word) | *(volatile const uint32_t *)(free_bits + word);
        active = active && !((w >> lane) & 1u);
      } else {
        const unsigned i = (unsigned)(fbase + 32 * v + lane);
        active = i < __ldg(n_unk + t);
        f = active ? __ldg(list + i) : 0;
        if (active) {
          const uint32_t w = *(volatile const uint32_t *)(free_bits + h.bits_off + (f >> 5));
          active = !((w >> (f & 31)) & 1u);          // another slice may have freed it meanwhile
        }
      }
      todo |= (active ? 1u : 0u) << v;
      const int fa = active ? f : 0;
      vf[v] = fa;
      const int ix = fa / yz, rem = fa - ix * yz;
      const int iy = rem / dZ;
      vx[v] = (float)ix;
      vy[v] = (float)iy;
      vz[v] = (float)(rem - iy * dZ);
    }
    unsigned steps = 0, found = 0u;
    const int k0 = (PHASE == 1) ? 0 : kPhase1Pairs + slice * kPairsPerItem;
    const int k1 = min(h.nact, k0 + (PHASE == 1 ? kPhase1Pairs : kPairsPerItem));
    const PairHot *tp = pairs + h.pairs_base;
    for (int k = k0; k < k1; ++k) {
      if (!__any_sync(0xffffffffu, todo != 0u)) break;           // every voxel of the item is settled
      const PairHot pc = load128(tp + k);
      if (k + 1 < k1) asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + k + 1));   // one 128-byte line
      const float *ub = ub_pool + 2 * pc.tab_off + 1;
      const uint16_t *lut = lut_pool + pc.tab_off * kLutPerRow;
      const float *ri_img = ri_pool + pc.ri_off;
      // materialise the three bases as 64-bit registers: per-test addresses are then ONE imad.wide each
      asm volatile("" : "+l"(ub), "+l"(lut), "+l"(ri_img));
      // all VPL tests are evaluated unconditionally (no divergence, their dependent chains interleave);
      // results of voxels this lane does not need are discarded
      int res[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) res[v] = fast_test(pc, vx[v], vy[v], vz[v], ub, lut, ri_img);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const bool need = (todo >> v) & 1u;
        res[v] = need ? ((pc.eps >= 0.f) ? res[v] : 1) : 0;
        steps += need ? 1u : 0u;
        if (res[v] == 2) {
          found |= 1u << v;
          todo &= ~(1u << v);
        }
        const unsigned umask = __ballot_sync(0xffffffffu, res[v] == 1);
        if (umask) {                                             // queue the undecided tests (warp-aggregated)
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(counter + 1, (unsigned long long)__popc(umask));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (res[v] == 1) {
            const unsigned long long slot = base + __popc(umask & ((1u << lane) - 1u));
            if (slot < (unsigned long long)queue_cap) {
              queue[slot] = make_int4(t, vf[v], pc.q, 0);
            } else if (exact_from_ids(t, vf[v], pc.q, L, vs, grids, trk_frame_off, poses, frame_sf, sensors,
                                      incl_pool, ri_pool)) {     // queue full: decide right here
              found |= 1u << v;
              todo &= ~(1u << v);
            }
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      if (PHASE == 1) {
        const unsigned fw = __ballot_sync(0xffffffffu, (found >> v) & 1u);
        if (fw && lane == 0) atomicOr(free_bits + h.bits_off + (int64_t)chunk * VPL + v, fw);
        if (h.nact > kPhase1Pairs) {                             // survivors go to the tracklet's phase-2 list
          const unsigned left = __ballot_sync(0xffffffffu, (todo >> v) & 1u);
          if (left) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(n_unk + t, (unsigned)__popc(left));
            base = __shfl_sync(0xffffffffu, base, 0);
            if ((todo >> v) & 1u) list[base + __popc(left & ((1u << lane) - 1u))] = vf[v];
          }
        }
      } else if ((found >> v) & 1u) {
        atomicOr(free_bits + h.bits_off + (vf[v] >> 5), 1u << (vf[v] & 31));
      }
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if (lane == 0 && n_steps && steps) atomicAdd((unsigned long long *)&n_steps[t], (unsigned long long)steps);
  }
}

__global__ void __launch_bounds__(256)
k_visibility_recheck(int L, const int64_t *__restrict__ trk_frame_off, const occb200_pose_t *__restrict__ poses,
                     const int32_t *__restrict__ frame_sf, const occb200_sensor_t *__restrict__ sensors,
                     const float *__restrict__ incl_pool, const float *__restrict__ ri_pool, double vs,
                     const TrkGrid *__restrict__ grids, const unsigned long long *__restrict__ counter,
                     const int4 *__restrict__ queue, long long queue_cap, uint32_t *__restrict__ free_bits,
                     int64_t *__restrict__ n_steps) {
  const long long n = min((long long)counter[1], queue_cap);
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int4 it = queue[k];
    const int t = it.x, f = it.y, q = it.z;
    const TrkGrid &g = grids[t];
    uint32_t *word = free_bits + g.bits_off + (f >> 5);
    const uint32_t bit = 1u << (f & 31);
    if (*(volatile uint32_t *)word & bit) continue;            // already proven free by another test
    double cx, cy, cz;
    voxel_centre(g, f, vs, cx, cy, cz);
    const int i = q / L, c = q - i * L;
    const int64_t f0 = trk_frame_off[t];
    const occb200_sensor_t *sn = sensors + (int64_t)__ldg(frame_sf + f0 + i) * L + c;
    if (exact_test(cx, cy, cz, poses + f0 + i, sn, incl_pool, ri_pool)) atomicOr(word, bit);
    if (n_steps) {                                             // one atomic per tracklet per warp, not per test
      const unsigned peers = __match_any_sync(__activemask(), t);
      if ((threadIdx.x & 31) == __ffs(peers) - 1)
        atomicAdd((unsigned long long *)&n_steps[t], (unsigned long long)__popc(peers));
    }
  }
}

// labels from the two bitsets: 1 occupied, 2 free, 0 unknown (occ_annotate.py:558-563); one CTA per tracklet
__global__ void __launch_bounds__(256)
k_labels(const TrkHot *__restrict__ hot, const TrkGrid *__restrict__ grids, const uint32_t *__restrict__ bits,
         const uint32_t *__restrict__ free_bits, int32_t *__restrict__ labels, int64_t *__restrict__ n_unknown,
         int32_t *__restrict__ status_out) {
  const int t = blockIdx.x;
  const TrkHot h = hot[t];
  int status = h.status;
  if (status == OCCB200_OK && grids[t].redo) {      // corrected tracklet: its flags are final only now
    const int fl = grids[t].flags;
    if (fl & 2) status = OCCB200_INDEX_ERROR;
    else if (!(fl & 1)) status = OCCB200_EMPTY_AFTER_FILTER;
    if (status != OCCB200_OK && threadIdx.x == 0 && blockIdx.y == 0) status_out[t] = status;
  }
  if (status != OCCB200_OK) return;
  int unk = 0;
  for (int f = blockIdx.y * blockDim.x + threadIdx.x; f < h.V; f += gridDim.y * blockDim.x) {
    const uint32_t o = bits[h.bits_off + (f >> 5)], fr = free_bits[h.bits_off + (f >> 5)];
    const bool occ = (o >> (f & 31)) & 1u;
    labels[h.label_off + f] = occ ? 1 : (((fr >> (f & 31)) & 1u) ? 2 : 0);
    unk += occ ? 0 : 1;
  }
  for (int o = 16; o > 0; o >>= 1) unk += __shfl_xor_sync(0xffffffffu, unk, o);
  if ((threadIdx.x & 31) == 0 && unk) atomicAdd((unsigned long long *)&n_unknown[t], (unsigned long long)unk);
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

// self-test hook: max |atan2_fast - atan2| over n pseudo-random f32 pairs (device-side check of kAtanErr)
__global__ void k_selftest_atan2(long long n, unsigned long long seed, unsigned long long *__restrict__ max_err) {
  double worst = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
    h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
    const float mag_x = exp2f((float)((h >> 8) & 31) - 20.f), mag_y = exp2f((float)((h >> 13) & 31) - 20.f);
    const float x = ((float)((h >> 20) & 0xfffff) / 524288.f - 1.f) * mag_x;
    const float y = ((float)((h >> 40) & 0xfffff) / 524288.f - 1.f) * mag_y;
    if (x == 0.f && y == 0.f) continue;
    const double err = fabs((double)atan2_fast(y, x) - atan2((double)y, (double)x));
    worst = fmax(worst, err);
  }
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(max_err, (unsigned long long)__double_as_longlong(worst));
}

}  // namespace occb200

using namespace occb200;

extern "C" int64_t occb200_annotate_workspace_bytes(int32_t T, int64_t F, int64_t total_label_slots, int64_t SF,
                                                    int32_t L, int64_t incl_len, int64_t pyr_tiles, int64_t items_cap) {
  return ws_layout(T, F, total_label_slots, SF, L, incl_len, pyr_tiles, items_cap, nullptr, nullptr);
}

// HOST helper: upper bound of the fast kernel's work items, from the host copies of label_off / trk_frame_off.
extern "C" int64_t occb200_annotate_items_cap(int32_t T, const int64_t *label_off, const int64_t *trk_frame_off,
                                              int32_t L) {
  int64_t n = 0;
  for (int t = 0; t < T; ++t) {
    const int64_t chunks = ceil_div(label_off[t + 1] - label_off[t], 32 * (kVPL1 < kVPL ? kVPL1 : kVPL));
    const int64_t slices = ceil_div((trk_frame_off[t + 1] - trk_frame_off[t]) * L, kPairsPerItem);
    n += chunks * slices;
  }
  return n;
}

extern "C" int64_t occb200_pyramid_tiles(int32_t H, int32_t W) {
  return (int64_t)((H + kTileR - 1) / kTileR) * ((W + kTileC - 1) / kTileC);
}

extern "C" int occb200_annotate_batch(const occb200_annotate_args_t *a, int64_t total, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(a != nullptr, "args is NULL");
  OCC_REQUIRE(a->T >= 0 && a->F >= 0 && a->L >= 1, "bad T/F/L");
  OCC_REQUIRE(a->SF >= 0 && a->incl_len >= 0 && a->pyr_tiles >= 0 && a->items_cap >= 0, "bad SF / incl_len / pyr_tiles / items_cap");
  OCC_REQUIRE(a->point_stride >= 3, "point_stride must be >= 3");
  OCC_REQUIRE(a->voxel_size > 0, "voxel_size must be positive");
  if (a->T == 0) return 0;
  Workspace w;
  const int64_t need = ws_layout(a->T, a->F, total, a->SF, a->L, a->incl_len, a->pyr_tiles, a->items_cap, (char *)a->workspace, &w);
  OCC_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= need, "workspace too small");
  const float vsf = (float)a->voxel_size;
  // shared-memory bitset of k_frame_voxelize: as many words as the largest label slot needs (unknown: the maximum)
  const int smem_words = (a->max_label_slots > 0)
                             ? (int)std::min<int64_t>(a->max_label_slots / 32 + 2, kSmemBitWords) : kSmemBitWords;
  const bool f64_only = (a->flags & 1) != 0;
  const int chunk = f64_only ? kChunk : kFastChunk;
  SideStream *side = nullptr;
  const bool fast = !f64_only && a->F > 0 && a->SF > 0;
  if (fast) {                                       // fork: sensor-side setup on the side stream
    if (side_stream(&side)) return 1;
    OCC_CUDA(cudaEventRecord(side->fork, stream));
    OCC_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    ProfScope ps(kProfPairSetup, side->stream);
    const int64_t n_sens = a->SF * a->L;
    OCC_CUDA(cudaMemsetAsync(w.tab_claim, 0, 4 * (size_t)std::max<int64_t>(w.incl_len, 1), side->stream));
    k_table_setup<<<(unsigned)n_sens, 256, 0, side->stream>>>(n_sens, a->sensors, a->incl_pool, w.sens, w.ub_pool,
                                                              w.lut_pool, w.tab_claim);
    OCC_KERNEL_OK("k_table_setup");
    k_pyr_scan<<<1, 1024, 0, side->stream>>>(n_sens, a->sensors, (a->flags & 2) ? (int64_t)-1 : w.pyr_tiles,
                                             w.pyr_off, w.pyr_flag);
    OCC_KERNEL_OK("k_pyr_scan");
    if (w.pyr_tiles > 0 && !(a->flags & 2)) {
      k_pyr_build<<<dim3((unsigned)n_sens, kPyrRowGroups), 256, 0, side->stream>>>(a->sensors, a->ri_pool, w.pyr_off,
                                                                                   w.pyr_flag, w.pyr);
      OCC_KERNEL_OK("k_pyr_build");
    }
    OCC_CUDA(cudaEventRecord(side->join, side->stream));
  }
  OCC_CUDA(cudaMemsetAsync(w.bits, 0, (char *)(w.free_bits + w.bits_words) - (char *)w.bits, stream));   // both bitsets
  {
    ProfScope ps(kProfInbox, stream);
    k_tracklet_presetup<<<(unsigned)ceil_div(a->T, 8), 256, 0, stream>>>(a->T, a->trk_frame_off, a->poses,
                                                                         a->frame_pt_off, a->label_off,
                                                                         vsf, chunk, w.grids, w.frame_trk, w.redo_count,
                                                                         w.counter, w.n_unk, a->n_unknown, a->n_steps);
    OCC_KERNEL_OK("k_tracklet_presetup");
  }
  if (a->F > 0) {
    ProfScope ps(kProfVoxelize, stream);
    k_frame_voxelize<<<(unsigned)a->F, kFrameThreads, 4 * smem_words, stream>>>(
        a->poses, a->points, a->point_stride, a->frame_pt_off, w.frame_kept, w.frame_trk, w.grids, w.bits, vsf, nullptr,
        nullptr, smem_words);
    OCC_KERNEL_OK("k_frame_voxelize");
  }
  {
    ProfScope ps(kProfSetup, stream);
    k_tracklet_setup<<<(unsigned)ceil_div(a->T, 8), 256, 0, stream>>>(
        a->T, a->trk_frame_off, a->poses, a->frame_pt_off, w.frame_kept, a->label_off, vsf, chunk, w.grids, w.bits,
        w.redo_list, w.redo_count, a->dims, a->sizes, a->status);
    OCC_KERNEL_OK("k_tracklet_setup");
    if (a->F > 0) {   // frames of corrected tracklets only (device-side list, usually short); in the fast path
                      // this runs on the side stream, next to k_pair_build, and joins before the ray-cast
      cudaStream_t rs = stream;
      if (fast) {
        OCC_CUDA(cudaStreamWaitEvent(stream, side->join, 0));      // the side stream is idle from here on
        OCC_CUDA(cudaEventRecord(side->fork, stream));
        OCC_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        rs = side->stream;
      }
      k_frame_voxelize<<<(unsigned)std::min<int64_t>(a->F, kNumSMs * 2), kFrameThreads, 4 * smem_words, rs>>>(
          a->poses, a->points, a->point_stride, a->frame_pt_off, w.frame_kept, w.frame_trk, w.grids, w.bits, vsf,
          w.redo_list, w.redo_count, smem_words);
      OCC_KERNEL_OK("k_frame_voxelize(redo)");
      if (fast) OCC_CUDA(cudaEventRecord(side->join, side->stream));
    }
  }
  if (f64_only) {
    ProfScope ps(kProfScan, stream);
    k_scan_chunks<<<1, 1024, 0, stream>>>(a->T, w.grids, w.chunk_off, w.counter);
    OCC_KERNEL_OK("k_scan_chunks");
  }
  const int64_t max_items = ceil_div(total, chunk) + a->T;
  if (f64_only) {
    const int grid = (int)std::min<int64_t>(max_items, (int64_t)kNumSMs * 8);
    ProfScope ps(kProfVisibility, stream);
    k_visibility_f64<<<grid, kChunk, 0, stream>>>(a->T, a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors,
                                                  a->incl_pool, a->ri_pool, a->voxel_size, a->label_off, w.grids,
                                                  w.chunk_off, w.counter, w.bits, a->labels, a->status,
                                                  a->n_unknown, a->n_steps);
    OCC_KERNEL_OK("k_visibility_f64");
    return 0;
  }
  if (fast) {
    ProfScope ps(kProfPairCull, stream);
    k_pair_build<<<(unsigned)a->T, 256, 0, stream>>>(a->T, a->L, a->trk_frame_off, a->label_off, w.grids, a->poses,
                                                     a->frame_sf, a->sensors, w.sens, a->voxel_size, w.pyr_off, w.pyr,
                                                     w.lut_pool, w.pyr_flag, w.pairs_c, w.hot, w.item_map,
                                                     (long long)w.items_cap, w.counter, a->status);
    OCC_KERNEL_OK("k_pair_build");
  }
  if (fast && a->F > 0) OCC_CUDA(cudaStreamWaitEvent(stream, side->join, 0));   // redo pass done: bits and flags final
  // flag bit 3 (tests): the kernels see a 64-entry recheck queue, so the decide-in-place path runs
  const long long queue_cap_used = (a->flags & 8) ? std::min<long long>(64, (long long)w.queue_cap) : (long long)w.queue_cap;
  {
    const int grid = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(w.items_cap, 1), kFastWarps),
                                            (int64_t)kNumSMs * OCC_MINB);
    ProfScope ps(kProfVisibility, stream);
    k_visibility_fast<1, kVPL1><<<grid, 32 * kFastWarps, 0, stream>>>(
        a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors, a->incl_pool, a->ri_pool, a->voxel_size, w.grids,
        w.counter, (long long)w.items_cap, w.bits, w.free_bits, w.item_map, w.hot, w.pairs_c, w.ub_pool,
        w.lut_pool, w.queue, queue_cap_used, w.unk_list, w.n_unk, a->n_steps);
    OCC_KERNEL_OK("k_visibility_fast<1>");
    k_phase2_emit<<<(unsigned)a->T, 256, 0, stream>>>(w.hot, w.n_unk, w.item_map, (long long)w.items_cap, w.counter);
    OCC_KERNEL_OK("k_phase2_emit");
    k_visibility_fast<2, kVPL><<<grid, 32 * kFastWarps, 0, stream>>>(
        a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors, a->incl_pool, a->ri_pool, a->voxel_size, w.grids,
        w.counter, (long long)w.items_cap, w.bits, w.free_bits, w.item_map, w.hot, w.pairs_c, w.ub_pool,
        w.lut_pool, w.queue, queue_cap_used, w.unk_list, w.n_unk, a->n_steps);
    OCC_KERNEL_OK("k_visibility_fast<2>");
  }
  {
    ProfScope ps(kProfRecheck, stream);
    k_visibility_recheck<<<kNumSMs * 4, 256, 0, stream>>>(a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors,
                                                          a->incl_pool, a->ri_pool, a->voxel_size, w.grids, w.counter,
                                                          w.queue, queue_cap_used, w.free_bits, a->n_steps);
    OCC_KERNEL_OK("k_visibility_recheck");
    k_labels<<<dim3((unsigned)a->T, 8), 256, 0, stream>>>(w.hot, w.grids, w.bits, w.free_bits, a->labels,
                                                          a->n_unknown, a->status);
    OCC_KERNEL_OK("k_labels");
  }
  return 0;
}

extern "C" int occb200_annotate_point_voxels(const occb200_annotate_args_t *a, int64_t total, float *loc_out,
                                             int32_t *q_out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(a != nullptr && loc_out != nullptr && q_out != nullptr, "NULL argument");
  OCC_REQUIRE(((uintptr_t)q_out & 15) == 0, "q_out must be 16-byte aligned");
  if (a->T == 0 || a->F == 0) return 0;
  Workspace w;
  ws_layout(a->T, a->F, total, a->SF, a->L, a->incl_len, a->pyr_tiles, a->items_cap, (char *)a->workspace, &w);
  k_frame_points<<<(unsigned)a->F, 256, 0, stream>>>(a->poses, a->points, a->point_stride, a->frame_pt_off,
                                                     w.frame_trk, w.grids, a->status, (float)a->voxel_size, loc_out,
                                                     q_out);
  OCC_KERNEL_OK("k_frame_points");
  return 0;
}

extern "C" int occb200_annotate_queue_stats(const occb200_annotate_args_t *a, int64_t total, int64_t *out_host,
                                            void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Workspace w;
  ws_layout(a->T, a->F, total, a->SF, a->L, a->incl_len, a->pyr_tiles, a->items_cap, (char *)a->workspace, &w);
  unsigned long long n = 0;
  OCC_CUDA(cudaMemcpyAsync(&n, w.counter + 1, 8, cudaMemcpyDeviceToHost, stream));
  OCC_CUDA(cudaStreamSynchronize(stream));
  out_host[0] = (int64_t)n;
  out_host[1] = w.queue_cap;
  return 0;
}

extern "C" void occb200_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
}

extern "C" int occb200_profile_kinds(void) { return kProfKinds; }

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

extern "C" int occb200_selftest_atan2(int64_t n, uint64_t seed, double *max_err_host, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned long long *d = nullptr;
  OCC_CUDA(cudaMallocAsync((void **)&d, 8, stream));
  OCC_CUDA(cudaMemsetAsync(d, 0, 8, stream));
  k_selftest_atan2<<<kNumSMs * 4, 256, 0, stream>>>(n, seed, d);
  OCC_KERNEL_OK("k_selftest_atan2");
  OCC_CUDA(cudaMemcpyAsync(max_err_host, d, 8, cudaMemcpyDeviceToHost, stream));
  OCC_CUDA(cudaFreeAsync(d, stream));
  OCC_CUDA(cudaStreamSynchronize(stream));
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
