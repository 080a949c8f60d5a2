// annotate.cu -- batched tracklet annotation: crop -> box frame -> voxelise -> visibility.
//
// Reference path: tools/occ/occ_annotate.py get_local_point_list :91-138 and
// OccAnnotator.annotate_trk :344-568 (normative step list: SURVEY.md Appendix A).
//
// Kernels (two streams with fork/join events, no host synchronisation; the whole call is capturable in a graph):
//   main:  memset -> k_crop_voxelize -> k_pair_build -> k_brick_cull -> k_visibility -> k_visibility_recheck -> k_labels
//   side:  k_pyr_build (range-image max pyramid, two levels; tiles flagged dead in args.ri_tile_live are skipped),
//          k_table_setup (row lookup tables); after k_pair_build: the re-voxelisation of the (rare) tracklets whose
//          optimistic grid was wrong, then k_brick_unknown (per-brick masks of the voxels without a point)
//
//   k_crop_voxelize   one CTA per equal share of the flat candidate-point array: optimistic grid of the tracklet,
//                     in-box test, box frame, quantise, bitset (A1/A2/A3)
//   k_tracklet_setup  (f64-only path; the fast path runs it as the prologue of k_pair_build) one warp per tracklet:
//                     box size = max over KEPT frames; rare re-voxelisation (A2/A3)
//   k_table_setup     one CTA per DISTINCT inclination table: row boundaries + 8-byte lookup cells
//   k_pyr_build       max of every 8x32 and 2x8 pixel tile of every range image
//   k_pair_build      one CTA per tracklet, one thread per (frame, LiDAR): voxel-index -> sensor-frame affine map
//                     (rotated so that the object sits on +x), pair-level cull, work items
//   k_brick_cull      one thread per (4x4x4-voxel brick, surviving pair): proves "no voxel of the brick can be
//                     free through this pair" from the fine pyramid level -> one mask bit
//   k_visibility      persistent warps over (brick, 16-pair slice) items: the range-image "ray-cast" in f32 with
//                     rigorous error margins; tests whose outcome is not certain are queued               (A4/A5)
//   k_visibility_recheck  the queued tests, re-evaluated with the reference's exact f64 arithmetic
//   k_labels          labels from the occupancy and free bitsets
//   k_visibility_f64  (flags bit 0) every test in exact f64 -- the slow, margin-free formulation
//
// Why the fast kernel is still bit-exact.  Per test the reference takes three discrete decisions from
// f64 quantities: the range-image row (nearest inclination), the column (rounded azimuth) and
// `range_image >= range`.  The fast kernel evaluates the same quantities in f32 from a per-(frame,
// LiDAR) affine map of the voxel index, with an explicit bound on |f32 value - reference f64 value|
// (derived at each use below).  A decision is accepted only if it stays the same anywhere inside that
// bound; otherwise the test goes to a queue and k_visibility_recheck redoes it with project_exact().
// Labels therefore never depend on f32 rounding; only the amount of rechecked work does.
// tools/experiments/fast_test_v2_emulate.py is a numpy f32 port of make_pair / k_table_setup / fast_test
// checked against the oracle's exact per-voxel decisions on the CPU.
#include <math.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "geom.cuh"

namespace occb200 {

constexpr int kChunk = 256;         // f64 kernel: voxels per work item == threads per CTA
constexpr int kFastWarps = 8;       // fast kernel: independent warps per CTA
#ifndef OCC_MINB
#define OCC_MINB 4
#endif
#ifndef OCC_PPI
#define OCC_PPI 16
#endif
#ifndef OCC_TAILSPLIT
#define OCC_TAILSPLIT 2
#endif
constexpr int kPairsPerItem = OCC_PPI;   // (frame, LiDAR) pairs one work item covers for its brick; divides 32
static_assert(32 % OCC_PPI == 0, "a slice of pairs must sit inside one 32-bit mask word");
constexpr int kTailSplit = OCC_TAILSPLIT; // tickets per item in the last round of k_visibility; divides kPairsPerItem
static_assert(OCC_PPI % OCC_TAILSPLIT == 0, "the tail split must divide the pairs of an item");
constexpr int kMaxSlices = 256;          // => at most 4096 pairs (frames x LiDARs) per tracklet on the fast path
constexpr float kAtanNarrow = 1.0e-6f;   // bound on |atan_poly(t) - atan(t)|, |t| <= 1 (derivation at atan_narrow)
constexpr float kAtanWide = 2.0e-6f;     // bound on |atan2_fast - atan2| (derivation at atan2_fast)
constexpr int kFrameStride = 8;          // pair order: frames 0,8,16,.. then 1,9,17,.. (spread viewpoints come first)
constexpr int kLutPerRow = 64;           // lookup-table cells reserved per inclination-table entry
constexpr int kTileR = 8, kTileC = 32;   // range-image max-pyramid tile (rows x columns): pair-level cull
constexpr int kFineR = 2, kFineC = 8;    // second pyramid level: brick-level cull; 16 fine tiles per coarse tile
constexpr int kBrick = 4;                // brick edge in voxels: 64 voxels = 2 per lane of one warp
#ifndef OCC_OVERLAP
#define OCC_OVERLAP 0  // 1: cull of slices >= 1 on a second side stream next to the visibility pass over slice 0 (measured
#endif                 // slower on C2, 94 vs 83 us for the ray-cast, and equal on 1 024 tracklets: kept for A/B only)
#ifndef OCC_VISDEBUG
#define OCC_VISDEBUG 0
#endif
#ifndef OCC_PP1
#define OCC_PP1 4      // pairs per loop iteration of the one-voxel-per-lane path (4 tests in flight per lane, like 2 x 2)
#endif
#ifndef OCC_PP2
#define OCC_PP2 2      // ... of the two-voxels-per-lane path
#endif
#ifndef OCC_BULK_STAGE
#define OCC_BULK_STAGE 0  // 1: k_visibility stages an item's pair records with ONE bulk copy (cp.async.bulk, the TMA engine's
#endif                    // 1-D form) issued by lane 0 as soon as the tracklet record is known, instead of 4 LDG.128 + 4 STS.128 per lane
#ifndef OCC_FT
#define OCC_FT 256
#endif
#ifndef OCC_FMINB
#define OCC_FMINB 4
#endif
constexpr int kFrameThreads = OCC_FT;
constexpr int kMaxGroup = 8;             // tracklet-frames one crop CTA handles at most
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
  int32_t redo;     // 1: the optimistic grid was wrong, the crop kernel runs again for this tracklet
  int32_t pad;
};

// Per DISTINCT inclination table (stored at the table's offset in incl_pool).  Row lookup happens in "u space":
// u(inc) = sin/(|sin| + cos), monotone with slope in [1/2, 1] over [-90, 90] degrees, so row boundaries never
// bunch up (unlike tan or sin).  Cell k of the lookup table is centred on u_k = (k - cell0m) / inv_w.
struct __align__(16) TabCoef {
  float inv_w;      // 1 / cell width; the cell width is half the closest pair of boundaries
  float cell0m;     // cell = rn(fma(u, inv_w, cell0m))
  float w;          // cell width
  int32_t ncell;
  int32_t H;
  int32_t ok;       // 0: no fast path through this table (not strictly descending, H < 2, too many cells ...)
  int32_t pad[2];
};
static_assert(sizeof(TabCoef) == 32, "TabCoef must be 32 bytes");

// One lookup cell: the row boundary inside the cell's extended interval [u_k - 0.75 w, u_k + 0.75 w] (at most
// one, the cells being half as wide as the closest pair of boundaries) and its index; without one, b = -4 and
// h = the number of boundaries above the cell.  row(u) = h + (b > u) for every u the cell can be selected for.
struct __align__(8) LutCell {
  float b;
  int32_t h;
};

// One (tracklet-frame, LiDAR) pair, everything one iteration of the visibility kernel needs in ONE 128-byte record
// (one L1 line, 8 x LDG.128 at a warp-uniform address).  Built by k_pair_build.
//   p' = bc + A (idx - cen):  sensor-frame position of voxel centre `idx`, ROTATED about the sensor's z axis by
//   -theta so that the centre of the tracklet's grid lies on the +x axis (y' ~ 0 there): azimuth = theta + atan(y'/x').
struct __align__(16) PairHot {
  float A[9];
  float bc[3];
  float eps;        // bound on the f32 position error (metres); < 0: no fast path for this pair (every test goes
                    // to the exact recheck)
  int32_t q;        // pair index inside the tracklet: frame * L + LiDAR
  float inv_w, cell0m;
  uint32_t ncm1;    // ncell - 1
  uint32_t last;    // H - 1
  float e15z;       // row margin: m = e15z / r + 1.5e-6
  float nkcol;      // -W / (2 pi)
  float c0f;        // colf_rel = fma(phi, nkcol, c0f): column relative to cint
  int32_t cint;     // first column (narrow pairs: + W where it is below W / 2, see fast_test)
  int32_t W;
  float ecolk, ecol;   // column margin: ecolk / rho + ecol
  float c1;         // range margin: m = r * 6.5e-7 + c1
  int32_t lut_off;  // first LutCell of the pair's table in lut_pool
  int32_t wide;     // 1: the object spans more than +-45 degrees of azimuth: full-quadrant arctangent
  int64_t ri_off;
  int32_t sens;     // sensor entry (pyramid offsets for the brick cull)
  int32_t pad;
};
static_assert(sizeof(PairHot) == 128, "PairHot must be 128 bytes");

// What the visibility kernels need of a tracklet, in one 64-byte record (4 x LDG.128).
struct __align__(16) TrkHot {
  int32_t V, dX, dY, dZ;
  int32_t status;       // final status (flags of the crop kernel folded in)
  int32_t nact;         // pairs that survived culling
  int64_t bits_off;
  int64_t pairs_base;   // first PairHot of the tracklet in pairs_c
  int64_t brick_base;   // first brick of the tracklet (global brick index)
  float cen[3];         // centre of the voxel-index lattice: (dims - 1) / 2
  int32_t pad;
};
static_assert(sizeof(TrkHot) == 64, "TrkHot must be 64 bytes");

struct Workspace {
  TrkGrid *grids;        // [T]
  int32_t *redo_list;    // [F] frames of tracklets whose optimistic grid was wrong
  int64_t *chunk_off;    // [T+1] (f64 path)
  // ---- zeroed by ONE memset at the start of every call
  char *zero_begin;
  unsigned long long *counter;   // [0] f64-path ticket, [1] recheck-queue length, [2] redo frames, [3], [4] visibility tickets,
                                 // [8 + 2s] brick items of slice s
  int32_t *trk_flags;    // [T] flags of the crop kernel (bit0 kept a point, bit1 index error)
  int32_t *frame_kept;   // [F] 1: the frame has an in-box point (set by the crop CTAs that see one)
  uint32_t *bits;        // occupancy bitsets, linear voxel order, tracklet t at label_off[t]/32 + t
  int64_t bits_words;
  uint32_t *free_brick;  // [2 * bricks] voxels proven free, BRICK order: bit j = lx*16 + ly*4 + lz of word pair 2*brick
  uint32_t *unk_brick;   // [2 * bricks] voxels inside the grid that hold no point, same order (k_brick_unknown; not zeroed)
  uint32_t *pair_mask;   // [mask_words * bricks] bit k: pair k of the tracklet cannot free any voxel of the brick
  TabCoef *tabcoef;      // [incl_len] (entries at table offsets only)
  char *zero_end;
  // ----
  float *ub_pool;        // [incl_len] u-space row boundaries of each table (scratch of k_table_setup)
  LutCell *lut_pool;     // [incl_len * kLutPerRow]
  int64_t incl_len;
  int4 *queue;           // recheck queue: (tracklet, voxel, pair index q = i*L + c, local brick * 64 + j)
  int64_t queue_cap;
  PairHot *pairs_c;      // [F*L] the non-culled pairs of each tracklet, compacted at trk_frame_off[t] * L
  TrkHot *hot;           // [T]
  int2 *item_map;        // [n_slices * bricks] brick items of slice s at s * bricks: (tracklet, bx | by << 10 | bz << 20)
  int64_t bricks;        // brick_off[T]
  int32_t mask_words;    // ceil(max_pairs / 32)
  int32_t n_slices;      // ceil(max_pairs / kPairsPerItem)
  float *pyr;            // [pyr_tiles] max of the range image over tiles of kTileR x kTileC pixels
  float *pyr2;           // [16 * pyr_tiles] second level, kFineR x kFineC pixels; image e at 16 * pyr_off[e],
                         // row length ceil(W / kFineC)
  int64_t pyr_tiles;
};

static int64_t ws_layout(int32_t T, int64_t F, int64_t total, int64_t SF, int32_t L, int64_t incl_len,
                         int64_t pyr_tiles, int64_t bricks, int32_t max_pairs, char *base, Workspace *w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const int32_t mask_words = (int32_t)std::max<int64_t>(ceil_div(std::max(max_pairs, 1), 32), 1);
  const int32_t n_slices = (int32_t)std::max<int64_t>(ceil_div(std::max(max_pairs, 1), kPairsPerItem), 1);
  int64_t o_grid = take(sizeof(TrkGrid) * (int64_t)T);
  int64_t o_redo = take(4 * F);
  int64_t o_choff = take(8 * ((int64_t)T + 1));
  // zeroed region
  int64_t o_cnt = take(8 * (8 + 2 * kMaxSlices));
  int64_t o_tf = take(4 * (int64_t)T);
  int64_t o_kept = take(4 * F);
  int64_t words = total / 32 + T + 1;
  int64_t o_bits = take(4 * words);
  int64_t o_fb = take(8 * std::max<int64_t>(bricks, 1));
  int64_t o_pm = take(4 * (int64_t)mask_words * std::max<int64_t>(bricks, 1));
  int64_t o_tc = take(sizeof(TabCoef) * std::max<int64_t>(incl_len, 1));
  int64_t o_zend = off;
  int64_t o_ub = take(4 * std::max<int64_t>(incl_len, 1));
  int64_t o_lut = take(sizeof(LutCell) * std::max<int64_t>(incl_len, 1) * kLutPerRow);
  // recheck queue: ~0.1 % of the EXECUTED tests are undecided in f32; room for 1/64 of the nominal tests, bounded.
  // Tests beyond the capacity are decided in place (exact_from_ids).
  const double nominal = (double)total * (T > 0 ? (double)F / T : 0.0) * L;
  int64_t qcap = (int64_t)std::min(std::max(nominal / 64.0, 65536.0), 16.0 * 1024 * 1024);
  int64_t o_q = take(16 * qcap);
  int64_t o_pc = take(sizeof(PairHot) * F * L);
  int64_t o_na = take(sizeof(TrkHot) * (int64_t)T);
  int64_t o_py = take(4 * pyr_tiles);
  int64_t o_py2 = take(4 * 16 * pyr_tiles);
  int64_t o_im = take(8 * (int64_t)n_slices * std::max<int64_t>(bricks, 1));
  int64_t o_ub2 = take(8 * std::max<int64_t>(bricks, 1));
  (void)SF;
  if (w) {
    w->grids = (TrkGrid *)(base + o_grid);
    w->frame_kept = (int32_t *)(base + o_kept);
    w->redo_list = (int32_t *)(base + o_redo);
    w->chunk_off = (int64_t *)(base + o_choff);
    w->zero_begin = base + o_cnt;
    w->counter = (unsigned long long *)(base + o_cnt);
    w->trk_flags = (int32_t *)(base + o_tf);
    w->bits = (uint32_t *)(base + o_bits);
    w->bits_words = words;
    w->free_brick = (uint32_t *)(base + o_fb);
    w->unk_brick = (uint32_t *)(base + o_ub2);
    w->pair_mask = (uint32_t *)(base + o_pm);
    w->tabcoef = (TabCoef *)(base + o_tc);
    w->zero_end = base + o_zend;
    w->ub_pool = (float *)(base + o_ub);
    w->lut_pool = (LutCell *)(base + o_lut);
    w->incl_len = incl_len;
    w->queue = (int4 *)(base + o_q);
    w->queue_cap = qcap;
    w->pairs_c = (PairHot *)(base + o_pc);
    w->hot = (TrkHot *)(base + o_na);
    w->item_map = (int2 *)(base + o_im);
    w->bricks = bricks;
    w->mask_words = mask_words;
    w->n_slices = n_slices;
    w->pyr = (float *)(base + o_py);
    w->pyr2 = (float *)(base + o_py2);
    w->pyr_tiles = pyr_tiles;
  }
  return off;
}

// ---------------------------------------------------------------------------------------------
// The sensor-side setup (row tables, range-image pyramid) does not depend on the tracklets, so it runs on
// a side stream concurrently with the crop/voxelise chain and joins before k_pair_build (fork/join with
// events; capturable in a CUDA graph).
struct SideStream {
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, fork2 = nullptr, join2 = nullptr;
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
    OCC_CUDA(cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
    OCC_CUDA(cudaEventCreateWithFlags(&s.fork2, cudaEventDisableTiming));
    OCC_CUDA(cudaEventCreateWithFlags(&s.join2, cudaEventDisableTiming));
  }
  *out = &s;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Optional per-kernel timing (bench.py's roofline): CUDA events recorded around each kernel of the
// pipeline on the launching stream; durations are summed per kernel when the profile is read.
enum { kProfCrop = 0, kProfSetup, kProfScan, kProfBrickCull, kProfVisibility, kProfSide, kProfRecheck, kProfPairBuild,
       kProfLabels, kProfKinds };
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
//   k_crop_voxelize      every CTA derives the tracklet's grid from the max over ALL frames that have candidate
//                        points, then ONE pass over the points of its frames: in-box test, box frame, quantise,
//                        set bits; records which frames had in-box points
//   k_tracklet_setup     recomputes the size from the kept frames; if it differs the tracklet's bits are
//                        cleared and a second k_crop_voxelize pass redoes just that tracklet
// Scalar divisions: the reference divides tensors by the python float voxel_size (:414-416, :425).  torch-CPU
// evaluates x / vs in IEEE f32; torch-CUDA evaluates x * (1 / vs) (its scalar-divisor fast path).  `inv_vs` == 0
// selects the CPU arithmetic (what the oracle restates), otherwise the product with inv_vs = 1.0f / vs (flag bit 5).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float div_vs(float x, float vsf, float inv_vs) {
  return inv_vs != 0.f ? __fmul_rn(x, inv_vs) : __fdiv_rn(x, vsf);
}

__device__ __forceinline__ void grid_from_size(TrkGrid &g, const float sz[3], float vsf, float inv_vs, int64_t cap,
                                               int chunk) {
  g.status = OCCB200_OK;
  for (int k = 0; k < 3; ++k) g.dims[k] = (int)ceilf(div_vs(sz[k], vsf, inv_vs));   // :414-416
  g.mb[0] = __fmul_rn(sz[0], -0.5f);           // min over corners of [0,0,0,w,l,h,0] (:422-423)
  g.mb[1] = __fmul_rn(sz[1], -0.5f);
  g.mb[2] = __fmul_rn(sz[2], 0.0f);
  g.V = (int64_t)g.dims[0] * g.dims[1] * g.dims[2];
  if (g.V > cap || g.V <= 0 || g.V >= (1ll << 31) || g.dims[0] >= 4096 || g.dims[1] >= 4096 || g.dims[2] >= 4096) {
    g.status = -1;                             // caller's slot too small: reported, nothing written
    g.V = 0;
  }
  g.nchunks = (int)((g.V + chunk - 1) / chunk);
}

// The optimistic / true grid of tracklet t from a box size; everything but `flags` and `redo`.
__device__ __forceinline__ TrkGrid make_grid(int t, int B, const float sz[3], float vsf, float inv_vs,
                                             const int64_t *__restrict__ label_off, int chunk) {
  TrkGrid g;
  g.B = B;
  g.flags = 0;
  g.nchunks = 0;
  g.V = 0;
  g.redo = 0;
  g.pad = 0;
  g.bits_off = label_off[t] / 32 + t;
  g.dims[0] = g.dims[1] = g.dims[2] = 0;
  g.mb[0] = g.mb[1] = g.mb[2] = 0.f;
  if (B < 10) g.status = OCCB200_SKIP_SHORT;                       // :344
  else if (sz[0] == -INFINITY) g.status = OCCB200_NO_POINTS;      // no candidate point at all (:129)
  else grid_from_size(g, sz, vsf, inv_vs, label_off[t + 1] - label_off[t], chunk);
  return g;
}

struct FramePose {     // what the crop needs of one tracklet-frame, in shared memory
  BoxTest bt;
  float c, s;          // torch f32 cos/sin(-yaw)
  float ox, oy, oz;
};

// One in-box point -> bit index in the tracklet's occupancy bitset, or -1; flags |= 1 kept, |= 2 index error,
// |= 4 the frame has an in-box point.
__device__ __forceinline__ int voxel_of_point(const FramePose &fp, const TrkGrid &g, float vsf, float inv_vs,
                                              float x, float y, float z, int &flags) {
  if (!pt_in_box(fp.bt, x, y, z)) return -1;
  flags |= 4;
  // local = (p + (-origin)) @ [[c,-s,0],[s,c,0],[0,0,1]]  (:117-122, lidar_box3d.py:165-184):
  // sgemm accumulates k = 0,1,2 as an FMA chain; the k=2 terms are exact no-ops.
  const float c = fp.c, s = fp.s;
  const float tx = __fadd_rn(x, -fp.ox), ty = __fadd_rn(y, -fp.oy), tz = __fadd_rn(z, -fp.oz);
  const float lx = __fmaf_rn(ty, s, __fmul_rn(tx, c));
  const float ly = __fmaf_rn(ty, c, __fmul_rn(tx, -s));
  const float lz = tz;
  // q = floor((local - min_bound) / vs)  (:425)
  float qx = floorf(div_vs(__fsub_rn(lx, g.mb[0]), vsf, inv_vs));
  float qy = floorf(div_vs(__fsub_rn(ly, g.mb[1]), vsf, inv_vs));
  float qz = floorf(div_vs(__fsub_rn(lz, g.mb[2]), vsf, inv_vs));
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
  return ((int)qx * g.dims[1] + (int)qy) * g.dims[2] + (int)qz;      // < V < 2^31 (grid_from_size)
}

// cuda_arith (flag bit 5): the in-box test as the reference's CUDA kernel evaluates it (device cosf / sinf, nvcc's
// FMA contraction) instead of the CPU kernel's arithmetic with host libm trig.
__device__ __forceinline__ BoxTest frame_box_test(const occb200_pose_t &ps, bool cuda_arith) {
  if (!cuda_arith) return make_box_test(ps.box, ps.cos_pib, ps.sin_pib, 0);
  const float a = box_rot_angle(ps.box[6]);
  return make_box_test(ps.box, cosf(a), sinf(a), 1);
}

__device__ __forceinline__ FramePose load_frame_pose(const occb200_pose_t &ps, bool cuda_arith) {
  FramePose fp;
  fp.bt = frame_box_test(ps, cuda_arith);
  fp.c = ps.cos_m;
  fp.s = ps.sin_m;
  fp.ox = ps.box[0];
  fp.oy = ps.box[1];
  fp.oz = ps.box[2];
  return fp;
}

#ifndef OCC_PPT
#define OCC_PPT 4
#endif
constexpr int kPtsPerThread = OCC_PPT;          // independent point loads in flight per thread
#ifndef OCC_CROP_ALIGN
#define OCC_CROP_ALIGN kFrameThreads   // (a whole iteration of the point loop, 1 024, left 16 % of the CTA slots of C2 empty)
#endif
constexpr int kCropChunkMin = OCC_CROP_ALIGN;                          // candidate points per crop CTA: a multiple of the CTA size,
constexpr int kCropChunkMax = 8 * kFrameThreads * kPtsPerThread;       // chosen per call so that the grid fills one resident wave
// First pass (redo_list == NULL): CTA c owns the candidate points [c * chunk, (c+1) * chunk) of the flat
// point array, whatever frames they belong to -- every CTA has the same amount of work (frames of a close object
// carry several times the points of a far one; one CTA per frame left the longest frame as the kernel's tail).
// Warp 0 finds the first frame of the chunk with a 32-way search of frame_pt_off (3 dependent loads for 32 768
// frames); the chunk is then processed in runs of consecutive frames of ONE tracklet (<= kMaxGroup).  For each run
// the CTA derives the tracklet's optimistic grid itself (max box size over the tracklet's frames that have candidate
// points: 2 dependent loads + one CTA reduction), zeroes a shared-memory bitset, streams the run's points as one
// flat range with kPtsPerThread loads in flight per thread, and ORs the bitset into the tracklet's global one.
// Second pass (redo_list != NULL): a small grid strides over the frames of the tracklets whose optimistic grid was
// wrong -- usually none -- whole frames, with the true grid from grids[].
__global__ void __launch_bounds__(kFrameThreads, OCC_FMINB)
k_crop_voxelize(int64_t F, int64_t P, int chunk, const occb200_pose_t *__restrict__ poses, const float *__restrict__ points,
                int stride, const int64_t *__restrict__ frame_pt_off, const int64_t *__restrict__ trk_frame_off,
                const int64_t *__restrict__ label_off, const int32_t *__restrict__ frame_trk,
                int32_t *__restrict__ frame_kept, int32_t *__restrict__ trk_flags, const TrkGrid *__restrict__ grids,
                uint32_t *__restrict__ bits, float vsf, float inv_vs, const int32_t *__restrict__ redo_list,
                const unsigned long long *__restrict__ redo_count, int smem_words) {
  extern __shared__ uint32_t s_bits[];            // smem_words words
  __shared__ FramePose s_fp[kMaxGroup];
  __shared__ int64_t s_off[kMaxGroup + 1];
  __shared__ __align__(16) int s_rel[kMaxGroup];      // first point of each frame of the run, relative to the run's first point in the chunk
  __shared__ unsigned s_keptmask;                 // bit k: frame k of the run has an in-box point
  __shared__ float s_red[kFrameThreads / 32][3];
  __shared__ TrkGrid s_grid;
  __shared__ int s_flags;
  __shared__ int64_t s_first;
  const bool redo_pass = redo_list != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n_work = redo_pass ? (long long)*redo_count : (long long)gridDim.x;
  for (long long wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    // the work item's point range [p0, p1) and first frame
    int64_t p0, p1, fa;
    if (redo_pass) {
      fa = (int64_t)redo_list[wi];
      p0 = frame_pt_off[fa];
      p1 = frame_pt_off[fa + 1];
    } else {
      p0 = (int64_t)wi * chunk;
      p1 = min(p0 + (int64_t)chunk, P);
      __syncthreads();
      if (warp == 0) {                              // largest f with frame_pt_off[f] <= p0 (p0 < P = frame_pt_off[F])
        int64_t lo = 0, hi = F;
        while (hi - lo > 1) {
          const int64_t step = (hi - lo + 31) / 32;
          const int64_t probe = lo + (int64_t)(lane + 1) * step;
          const bool le = probe < hi && frame_pt_off[probe] <= p0;
          const int c = __popc(__ballot_sync(0xffffffffu, le));       // true for a prefix of the lanes
          hi = min(hi, lo + (int64_t)(c + 1) * step);
          lo = lo + (int64_t)c * step;
        }
        if (lane == 0) s_first = lo;
      }
      __syncthreads();
      fa = s_first;
    }
    int t_grid = -1;                                // tracklet whose grid s_grid holds
    while (fa < F && frame_pt_off[fa] < p1) {       // one run = consecutive frames of ONE tracklet
      __syncthreads();                              // the previous run's shared state is no longer read
      const int t = frame_trk[fa];
      const int64_t f0 = trk_frame_off[t], f1 = trk_frame_off[t + 1];
      int64_t fb = redo_pass ? fa + 1 : min(f1, fa + (int64_t)kMaxGroup);
      if (!redo_pass) {                             // frames that start at or after the chunk's end belong to the next CTA
        int c = 0;
        for (int i = 1; i < (int)(fb - fa); ++i) c += (frame_pt_off[fa + i] < p1) ? 1 : 0;   // offsets are non-decreasing
        fb = fa + 1 + c;
      }
      const int nfr = (int)(fb - fa);
      if (redo_pass) {
        if (threadIdx.x == 0) s_grid = grids[t];
      } else if (t != t_grid) {                     // a chunk rarely leaves its tracklet: the grid is derived once
        t_grid = t;
        float sz[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int64_t f = f0 + threadIdx.x; f < f1; f += kFrameThreads)
          if (frame_pt_off[f + 1] > frame_pt_off[f])   // a frame without candidates cannot be a kept frame
#pragma unroll
            for (int k = 0; k < 3; ++k) sz[k] = fmaxf(sz[k], poses[f].box[3 + k]);
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int k = 0; k < 3; ++k) sz[k] = fmaxf(sz[k], __shfl_xor_sync(0xffffffffu, sz[k], o));
        if (lane == 0)
#pragma unroll
          for (int k = 0; k < 3; ++k) s_red[warp][k] = sz[k];
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            for (int w2 = 1; w2 < kFrameThreads / 32; ++w2) sz[k] = fmaxf(sz[k], s_red[w2][k]);
          s_grid = make_grid(t, (int)(f1 - f0), sz, vsf, inv_vs, label_off, 32);
        }
      }
      if (threadIdx.x < nfr) {
        s_fp[threadIdx.x] = load_frame_pose(poses[fa + threadIdx.x], inv_vs != 0.f);
        s_off[threadIdx.x] = frame_pt_off[fa + threadIdx.x];
      }
      if (threadIdx.x < kMaxGroup) {                // 32-bit offsets for the per-point frame search (INT_MAX: no such frame)
        const int64_t first = redo_pass ? frame_pt_off[fa] : max(frame_pt_off[fa], p0);
        const int64_t rel = (threadIdx.x < nfr) ? frame_pt_off[fa + threadIdx.x] - first : (int64_t)INT_MAX;
        s_rel[threadIdx.x] = (int)min(max(rel, (int64_t)INT_MIN), (int64_t)INT_MAX);
      }
      if (threadIdx.x == 0) {
        s_off[nfr] = frame_pt_off[fb];
        s_flags = 0;
        s_keptmask = 0u;
      }
      __syncthreads();
      const TrkGrid g = s_grid;
      if (g.status != OCCB200_OK) {
        fa = fb;
        continue;
      }
      const int words = (int)((g.V + 31) / 32);
      const bool use_smem = words <= smem_words;
      uint32_t *gbits = bits + g.bits_off;
      if (use_smem)
        for (int w2 = threadIdx.x; w2 < words; w2 += kFrameThreads) s_bits[w2] = 0u;
      __syncthreads();

      const int64_t n0 = max(s_off[0], p0), n1 = min(s_off[nfr], p1);    // the run's points inside the chunk
      int flags = 0;
      unsigned kept = 0u;                                                 // frames of the run this thread saw an in-box point in
      // the run is walked in pieces of < 2^30 points (one piece, as a rule) with 32-bit point numbers relative to n0
      const float *prun = points + n0 * stride;
      const int4 rel_a = *reinterpret_cast<const int4 *>(s_rel), rel_b = *reinterpret_cast<const int4 *>(s_rel + 4);
      static_assert(kMaxGroup == 8, "the frame search reads eight offsets");
      for (int64_t piece = 0; piece < n1 - n0; piece += (1 << 30)) {
      const int npts = (int)min(n1 - n0 - piece, (int64_t)(1 << 30));
      const int jr0 = (int)piece;                                         // (pieces beyond the first: redo pass only, where nfr == 1)
      for (int base = 0; base < npts; base += kFrameThreads * kPtsPerThread) {   // warp-uniform trip count
        float px[kPtsPerThread], py[kPtsPerThread], pz[kPtsPerThread];
#pragma unroll
        for (int u = 0; u < kPtsPerThread; ++u) {
          const int j = base + u * kFrameThreads + (int)threadIdx.x;
          const float *p = prun + (piece + j) * stride;
          const bool ok = j < npts;
          px[u] = ok ? ld_stream(p) : 0.f;
          py[u] = ok ? ld_stream(p + 1) : 0.f;
          pz[u] = ok ? ld_stream(p + 2) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kPtsPerThread; ++u) {
          const int j = base + u * kFrameThreads + (int)threadIdx.x;
          int idx = -1;
          if (j < npts) {
            const int jr = jr0 + j;
            // frame of the point inside the run: the number of frames that start at or before it
            // (sign bit of rel - 1 - jr: set iff jr >= rel; both lie in [0, 2^31), no overflow)
            const int nj = ~jr;
            const unsigned k = ((unsigned)(rel_a.y + nj) >> 31) + ((unsigned)(rel_a.z + nj) >> 31) +
                               ((unsigned)(rel_a.w + nj) >> 31) + ((unsigned)(rel_b.x + nj) >> 31) +
                               ((unsigned)(rel_b.y + nj) >> 31) + ((unsigned)(rel_b.z + nj) >> 31) +
                               ((unsigned)(rel_b.w + nj) >> 31);
            int fl = 0;
            idx = voxel_of_point(s_fp[k], g, vsf, inv_vs, px[u], py[u], pz[u], fl);
            kept |= (fl & 4) ? (1u << k) : 0u;
            flags |= fl & 3;
          }
          int word = -1;
          uint32_t bit = 0u;
          if (idx >= 0) {
            word = idx >> 5;
            bit = 1u << (idx & 31);
            // most points land in voxels that are already marked: look before touching an atomic
            const uint32_t cur = use_smem ? s_bits[word] : __ldg(gbits + word);
            if (cur & bit) word = -1;
          }
          if (use_smem) {
            // shared-memory bitset: a plain atomic per NEW bit (the look above has removed most points; a warp-level
            // match_any dedup in front of it cost ~80 instructions per point -- its loop runs once per distinct word)
            if (word >= 0) atomicOr(&s_bits[word], bit);
          } else if (__any_sync(0xffffffffu, word >= 0)) {
            // global bitset (grids beyond the shared-memory limit): lanes on the same word merge their bits first
            const unsigned peers = __match_any_sync(0xffffffffu, word);
            const uint32_t merged = __reduce_or_sync(peers, word >= 0 ? bit : 0u);
            if (word >= 0 && lane == __ffs(peers) - 1) atomicOr(&gbits[word], merged);
          }
        }
      }
      }
      if (flags) atomicOr(&s_flags, flags);
      if (kept) atomicOr(&s_keptmask, kept);
      __syncthreads();
      if (use_smem)
        for (int w2 = threadIdx.x; w2 < words; w2 += kFrameThreads) {
          const uint32_t v = s_bits[w2];
          if (v) atomicOr(&gbits[w2], v);
        }
      // a frame may be shared with the neighbouring chunks: frame_kept starts at 0 and is only ever set
      if (threadIdx.x < nfr && !redo_pass && ((s_keptmask >> threadIdx.x) & 1u)) frame_kept[fa + threadIdx.x] = 1;
      if (threadIdx.x == 0 && s_flags) atomicOr(&trk_flags[t], s_flags);
      fa = fb;
    }
  }
}

// --save-mean-var support (occ_annotate.py:627-645): for every candidate point, its box-frame coordinates and raw
// quantised voxel coordinates with the tracklet's FINAL grid, exactly as k_crop_voxelize derives them.  Row
// (t, qx, qy, qz) for a point the reference keeps (in the box, q < dims), (-1, 0, 0, 0) otherwise.  Negative
// coordinates are NOT wrapped here: the reference groups by the raw values (sst_ops.py:150-181).
__global__ void __launch_bounds__(256)
k_frame_points(const occb200_pose_t *__restrict__ poses, const float *__restrict__ points, int stride,
               const int64_t *__restrict__ frame_pt_off, const int32_t *__restrict__ frame_trk,
               const TrkGrid *__restrict__ grids, const int32_t *__restrict__ status, float vsf, float inv_vs,
               float *__restrict__ loc_out, int32_t *__restrict__ q_out) {
  const int64_t f = blockIdx.x;
  const int t = frame_trk[f];
  const TrkGrid g = grids[t];
  const bool live = status[t] == OCCB200_OK;
  const occb200_pose_t ps = poses[f];
  const BoxTest bt = frame_box_test(ps, inv_vs != 0.f);
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
      const float qx = floorf(div_vs(__fsub_rn(lx, g.mb[0]), vsf, inv_vs));
      const float qy = floorf(div_vs(__fsub_rn(ly, g.mb[1]), vsf, inv_vs));
      const float qz = floorf(div_vs(__fsub_rn(lz, g.mb[2]), vsf, inv_vs));
      if (qx < dX && qy < dY && qz < dZ) row = make_int4(t, (int)qx, (int)qy, (int)qz);
    }
    loc_out[3 * j + 0] = lx;
    loc_out[3 * j + 1] = ly;
    loc_out[3 * j + 2] = lz;
    reinterpret_cast<int4 *>(q_out)[j] = row;
  }
}

// One warp per tracklet: the grid the crop kernel assumed (size = max over the frames with candidate points) against
// the true one (max over KEPT frames, occ_annotate.py:111-112, :132-133, box_mode="max"); writes grids[t], dims,
// sizes, the first status, and zeroes the per-tracklet counters of the call.  Returns the tracklet's grid (every lane).
__device__ __forceinline__ TrkGrid
tracklet_setup_warp(int t, int lane, const int64_t *__restrict__ trk_frame_off, const occb200_pose_t *__restrict__ poses,
                    const int64_t *__restrict__ frame_pt_off, const int32_t *__restrict__ frame_kept,
                    const int64_t *__restrict__ label_off, float vsf, float inv_vs, int chunk,
                    TrkGrid *__restrict__ grids, uint32_t *__restrict__ bits, int32_t *__restrict__ trk_flags,
                    int32_t *__restrict__ redo_list, unsigned long long *__restrict__ redo_count,
                    int32_t *__restrict__ dims_out, float *__restrict__ sizes_out, int32_t *__restrict__ status_out,
                    int64_t *__restrict__ n_unknown, int64_t *__restrict__ n_steps) {
  const int64_t f0 = trk_frame_off[t], f1 = trk_frame_off[t + 1];
  float sz[3] = {-INFINITY, -INFINITY, -INFINITY}, sz_all[3] = {-INFINITY, -INFINITY, -INFINITY};
  int kept = 0;
  for (int64_t f = f0 + lane; f < f1; f += 32) {
    const bool has = frame_pt_off[f + 1] > frame_pt_off[f];
    const bool k_ = has && frame_kept[f] != 0;
    kept += k_ ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float b = poses[f].box[3 + k];
      if (has) sz_all[k] = fmaxf(sz_all[k], b);         // what the crop kernel assumed
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
  TrkGrid g = make_grid(t, (int)(f1 - f0), sz_all, vsf, inv_vs, label_off, chunk);   // the optimistic grid
  g.flags = trk_flags[t];
  if (g.status == OCCB200_OK) {                   // (-1, slot too small for the optimistic grid, stays -1)
    if (kept == 0) {
      g.status = OCCB200_NO_POINTS;               // :129
      g.V = 0;
      g.nchunks = 0;
      g.dims[0] = g.dims[1] = g.dims[2] = 0;
    } else if (sz[0] != sz_all[0] || sz[1] != sz_all[1] || sz[2] != sz_all[2]) {
      // a frame without in-box points carried the largest box: clear the bits set with the optimistic grid
      // and let the second crop pass redo this tracklet with the true one.
      const int old_words = (int)((g.V + 31) / 32);
      grid_from_size(g, sz, vsf, inv_vs, label_off[t + 1] - label_off[t], chunk);
      g.flags = 0;
      g.redo = 1;
      uint32_t *gbits = bits + g.bits_off;
      for (int w = lane; w < old_words; w += 32) gbits[w] = 0u;
      if (lane == 0) trk_flags[t] = 0;
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(redo_count, (unsigned long long)kept);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int64_t f = f0 + lane; f - lane < f1; f += 32) {        // kept frames, warp-compacted
        const bool k_ = f < f1 && frame_pt_off[f + 1] > frame_pt_off[f] && frame_kept[f] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k_);
        if (k_) redo_list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)f;
        base += __popc(m);
      }
    }
  }
  if (lane == 0) {
    grids[t] = g;
    for (int k = 0; k < 3; ++k) {
      dims_out[3 * t + k] = g.dims[k];
      sizes_out[3 * t + k] = (g.status == OCCB200_OK) ? sz[k] : 0.f;
    }
    status_out[t] = g.status;                    // refined by k_pair_build / k_labels (flags)
    n_unknown[t] = 0;
    if (n_steps) n_steps[t] = 0;
  }
  return g;
}

// The set-up as a kernel of its own: the all-f64 path (flag bit 0); the fast path runs it as k_pair_build's prologue.
__global__ void __launch_bounds__(256)
k_tracklet_setup(int T, const int64_t *__restrict__ trk_frame_off, const occb200_pose_t *__restrict__ poses,
                 const int64_t *__restrict__ frame_pt_off, const int32_t *__restrict__ frame_kept,
                 const int64_t *__restrict__ label_off, float vsf, float inv_vs, int chunk,
                 TrkGrid *__restrict__ grids, uint32_t *__restrict__ bits, int32_t *__restrict__ trk_flags,
                 int32_t *__restrict__ redo_list, unsigned long long *__restrict__ redo_count,
                 int32_t *__restrict__ dims_out, float *__restrict__ sizes_out, int32_t *__restrict__ status_out,
                 int64_t *__restrict__ n_unknown, int64_t *__restrict__ n_steps) {
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per tracklet
  if (t >= T) return;
  tracklet_setup_warp(t, threadIdx.x & 31, trk_frame_off, poses, frame_pt_off, frame_kept, label_off, vsf, inv_vs, chunk,
                      grids, bits, trk_flags, redo_list, redo_count, dims_out, sizes_out, status_out, n_unknown, n_steps);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_scan_chunks(int T, const TrkGrid *__restrict__ grids, int64_t *__restrict__ chunk_off) {
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
                 const int32_t *__restrict__ trk_flags, int32_t *__restrict__ labels,
                 uint8_t *__restrict__ labels_u8, int32_t *__restrict__ status_out,
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
    // status refinement once per tracklet (flags are final: both crop passes have completed)
    int status = g.status;
    if (status == OCCB200_OK) {
      const int fl = trk_flags[t];
      if (fl & 2) status = OCCB200_INDEX_ERROR;
      else if (!(fl & 1)) status = OCCB200_EMPTY_AFTER_FILTER;
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
    if (active) {                                  // :558-563
      const int lab = occupied ? 1 : (is_free ? 2 : 0);
      if (labels) labels[label_off[t] + f] = lab;
      if (labels_u8) labels_u8[label_off[t] + f] = (uint8_t)lab;
    }
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
// midpoints m_h = (t_h + t_{h+1})/2 that lie above inc.  Boundaries are kept as ub_h = u(m_h).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double u_of_angle(double a) {
  const double s = sin(a), c = cos(a);
  return s / (fabs(s) + c);
}

// kTabSplit CTAs per DISTINCT table (the frames of a segment share one table per LiDAR): every CTA derives the
// boundaries and the cell geometry (a few hundred values), CTA y writes the lookup cells k = y*256 + tid (+ stride):
// one binary search per thread instead of sixteen in sequence.
constexpr int kTabSplit = 16;
__global__ void __launch_bounds__(256)
k_table_setup(int n_tables, const int64_t *__restrict__ table_off, const int32_t *__restrict__ table_H,
              const float *__restrict__ incl_pool, TabCoef *__restrict__ tabcoef, float *__restrict__ ub_pool,
              LutCell *__restrict__ lut_pool) {
  __shared__ float s_min[256];
  __shared__ TabCoef s_tc;
  extern __shared__ float s_ub[];                 // H - 1 boundaries (this CTA's private copy)
  const int e = blockIdx.x;
  if (e >= n_tables) return;
  const int H = table_H[e];
  const int64_t off = table_off[e];
  const float *tab = incl_pool + off;
  LutCell *lut = lut_pool + off * kLutPerRow;
  (void)ub_pool;
  const bool candidate = H >= 2 && H < 8192 && off < (1ll << 24);
  float local_min = INFINITY;
  if (candidate)
    for (int h = threadIdx.x; h < H - 1; h += blockDim.x) {
      const double m = 0.5 * ((double)tab[h] + (double)tab[h + 1]);
      s_ub[h] = (float)u_of_angle(m);
    }
  __syncthreads();
  const float *ub = s_ub;
  if (candidate) {
    for (int h = threadIdx.x; h < H - 2; h += blockDim.x) local_min = fminf(local_min, ub[h] - ub[h + 1]);
    // the table must stay inside (-90, 90) degrees for u() to be monotone, and descend (tab[h] > tab[h+1]: a
    // non-positive spacing fails the test below; NaN entries fail this one)
    for (int h = threadIdx.x; h < H; h += blockDim.x)
      if (!(fabsf(tab[h]) < 1.5707f)) local_min = -1.f;
    if (H == 2 && threadIdx.x == 0 && !(tab[0] > tab[1])) local_min = -1.f;
  }
  s_min[threadIdx.x] = local_min;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) s_min[threadIdx.x] = fminf(s_min[threadIdx.x], s_min[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    TabCoef tc;
    tc.inv_w = 0.f; tc.cell0m = 0.f; tc.w = 0.f; tc.ncell = 0; tc.H = H; tc.ok = 0; tc.pad[0] = tc.pad[1] = 0;
    float spacing = s_min[0];
    if (candidate && H == 2 && spacing > 0.f) spacing = 0.25f;
    if (candidate && spacing > 1e-6f && isfinite(spacing)) {
      // cells half as wide as the closest pair of boundaries, centred on u_k = (k - cell0m) / inv_w, covering
      // [-1.02, 1.02]
      const float w = 0.5f * spacing;
      const int ncell = (int)ceilf(2.04f / w) + 2;
      if (ncell <= H * kLutPerRow) {
        tc.w = w;
        tc.inv_w = 1.0f / w;
        tc.cell0m = 1.02f * tc.inv_w;
        tc.ncell = ncell;
        tc.ok = 1;
      }
    }
    s_tc = tc;
    if (blockIdx.y == 0) tabcoef[off] = tc;
  }
  __syncthreads();
  const TabCoef tc = s_tc;
  if (!tc.ok) return;
  const double w = (double)tc.w, inv_w = (double)tc.inv_w, c0 = (double)tc.cell0m;
  for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < tc.ncell; k += gridDim.y * blockDim.x) {
    const double uk = ((double)k - c0) / inv_w;
    int lo = 0, hi = H - 1;                       // count of ub_h > uk (ub descending)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((double)ub[mid] > uk) lo = mid + 1; else hi = mid;
    }
    // the boundary nearest to uk is ub[lo - 1] (just above) or ub[lo] (at or below); at most one of them can sit in
    // the extended cell [uk - 0.75 w, uk + 0.75 w]: the boundaries are >= 2 w apart
    LutCell c;
    c.b = -4.f;
    c.h = lo;
    if (lo >= 1 && (double)ub[lo - 1] - uk <= 0.75 * w) { c.b = ub[lo - 1]; c.h = lo - 1; }
    else if (lo < H - 1 && uk - (double)ub[lo] <= 0.75 * w) { c.b = ub[lo]; c.h = lo; }
    lut[k] = c;
  }
}

// row(u) = number of boundaries above u, for ANY u (the caller supplies the error padding): used by the culls.
__device__ __forceinline__ int row_of_u(const LutCell *__restrict__ lut, float inv_w, float cell0m, int ncell,
                                        float u) {
  const int cell = max(0, min((int)rintf(fmaf(u, inv_w, cell0m)), ncell - 1));
  const LutCell c = lut[cell];
  return c.h + ((c.b > u) ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------
// Range-image max pyramid: pyr[tile] = max of the image over kTileR x kTileC pixels, pyr2 over kFineR x kFineC.
// Used only to PROVE that a (frame, LiDAR) pair cannot free any voxel of a tracklet / a brick (all returns in
// the window it projects to are nearer than its nearest voxel), so that the pair is skipped.
// ---------------------------------------------------------------------------------------------
constexpr int kPyrRowGroups = 8;
__global__ void __launch_bounds__(256)
k_pyr_build(const occb200_sensor_t *__restrict__ sensors, const float *__restrict__ ri_pool,
            const int64_t *__restrict__ pyr_off, float *__restrict__ pyr, float *__restrict__ pyr2,
            const uint8_t *__restrict__ tile_live) {
  // tile_live (optional, args.ri_tile_live): one byte per 8x32 tile, 0 = no visibility test of the batch can read a
  // pixel of the tile.  Such tiles are skipped -- both levels were zeroed by the caller: their maxima count as 0,
  // which is exact for every pixel a test can read (ri_windows.cu: "why the labels cannot change").
  const int e = blockIdx.x;                        // sensor entry; blockIdx.y = row group
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const occb200_sensor_t &sn = sensors[e];
  const int H = sn.H, W = sn.W;
  const int ntr = (H + kTileR - 1) / kTileR, ntc = (W + kTileC - 1) / kTileC;
  const float *img = ri_pool + sn.ri_off;
  float *out = pyr + pyr_off[e];
  const uint8_t *live = tile_live ? tile_live + pyr_off[e] : nullptr;
  float *out2 = pyr2 + 16 * pyr_off[e];            // fine level: 16 slots per coarse tile; row pitch 4 * ntc tiles
                                                   // (>= ceil(W / 8), a multiple of 4: rows are read with 16-byte loads)
  const int nr2 = (H + kFineR - 1) / kFineR, nc2 = (W + kFineC - 1) / kFineC, pitch2 = 4 * ntc;
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
      bool on[kP];
      bool any_on = false;
#pragma unroll
      for (int u = 0; u < kP; ++u) {               // (warp-uniform) a pair is read if one of its two tiles is live
        const int p = i0 + u;
        const int tr = p / npc, tp = p - tr * npc;
        on[u] = p < npair;
        if (live && on[u]) on[u] = live[tr * ntc + 2 * tp] || (2 * tp + 1 < ntc && live[tr * ntc + 2 * tp + 1]);
        any_on = any_on || on[u];
      }
      if (!any_on) continue;
#pragma unroll
      for (int u = 0; u < kP; ++u) {
        const int p = i0 + u;
        const int tr = p / npc, tp = p - tr * npc;
        const int col = tp * (2 * kTileC) + 2 * lane;
#pragma unroll
        for (int r = 0; r < kTileR; ++r) {
          const int row = tr * kTileR + r;
          v[u][r] = (on[u] && col < W && row < H)
                        ? ld_stream2(reinterpret_cast<const float2 *>(img + (int64_t)row * W + col))
                        : make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < kP; ++u) {
        if (!on[u]) continue;
        const int p = i0 + u;
        const int tr = p / npc, tp = p - tr * npc;
        float m = 0.f;
#pragma unroll
        for (int r = 0; r < kTileR; ++r) m = fmaxf(m, fmaxf(v[u][r].x, v[u][r].y));
        // fine level: 2 rows x 8 columns = rows (2j, 2j+1) of the four lanes that share columns [8g, 8g+8)
#pragma unroll
        for (int j = 0; j < kTileR / kFineR; ++j) {
          float m2 = fmaxf(fmaxf(v[u][2 * j].x, v[u][2 * j].y), fmaxf(v[u][2 * j + 1].x, v[u][2 * j + 1].y));
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
          const int r2 = tr * (kTileR / kFineR) + j, c2 = tp * (2 * kTileC / kFineC) + (lane >> 2);
          if ((lane & 3) == 0 && p < npair && r2 < nr2 && c2 < nc2) out2[r2 * pitch2 + c2] = m2;
        }
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));   // within 16 lanes
        const int tc = 2 * tp + (lane >> 4);       // lanes 0-15: first tile of the pair, 16-31: second
        if ((lane & 15) == 0 && p < npair && tc < ntc) out[tr * ntc + tc] = m;
      }
    }
    return;
  }
  for (int i0 = (blockIdx.y * 8 + warp) * kU; i0 < ntile; i0 += kPyrRowGroups * 8 * kU) {
    float v[kU][kTileR];
    bool on[kU];
    bool any_on = false;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      on[u] = i0 + u < ntile && (!live || live[i0 + u]);
      any_on = any_on || on[u];
    }
    if (!any_on) continue;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int tile = i0 + u;
      const int tr = tile / ntc, tc = tile - tr * ntc;
      const int col = tc * kTileC + lane;
#pragma unroll
      for (int r = 0; r < kTileR; ++r) {           // range images are >= 0 (0 = no return)
        const int row = tr * kTileR + r;
        v[u][r] = (on[u] && col < W && row < H) ? ld_stream(img + (int64_t)row * W + col) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!on[u]) continue;
      float m = 0.f;
#pragma unroll
      for (int r = 0; r < kTileR; ++r) m = fmaxf(m, v[u][r]);
      {                                            // fine level: rows (2j, 2j+1) of the eight lanes of a column group
        const int tile = i0 + u;
        const int tr = tile / ntc, tc = tile - tr * ntc;
#pragma unroll
        for (int j = 0; j < kTileR / kFineR; ++j) {
          float m2 = fmaxf(v[u][2 * j], v[u][2 * j + 1]);
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 4));
          const int r2 = tr * (kTileR / kFineR) + j, c2 = tc * (kTileC / kFineC) + (lane >> 3);
          if ((lane & 7) == 0 && tile < ntile && r2 < nr2 && c2 < nc2) out2[r2 * pitch2 + c2] = m2;
        }
      }
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
// The record stores the map ROTATED about the sensor's z axis by -theta, theta = azimuth of the centre of the
// index lattice, and relative to that centre:  p' = bc + A' (idx - cen),  bc = Rz(-theta) p(cen) = (rho_c, ~0, z_c).
// Then azimuth(p) = theta + atan2(y', x') with |y'/x'| small, and the f32 rounding of y' happens at the magnitude
// of the OBJECT (metres), not of its distance (tens of metres).
//
// Error bounds (per component; 2^-24 = half an ulp; factor 6 = 3 FMA roundings + coefficient roundings + slack):
//   eps_y  = 6 * 2^-24 * (|bc_y| + sum_k |A'_yk| span_k / 2)          -- bc_y ~ 0
//   eps_xz = 6 * 2^-24 * max over x, z of (|bc_r| + sum_k |A'_rk| span_k / 2)
// the reference's own f64 roundings (~1e-14 m) vanish in the slack.
//
// Pair-level culling.  All voxel centres lie in the ball (centre pc = p(cen), radius R = half the lattice
// diagonal).  Seen from the sensor the ball spans inclinations inc_c +- asin(R/d) and azimuths
// az_c +- asin(R/rho_c); the reference's row/column rules are monotone in those angles, so every
// pixel any centre can map to lies in the row/column window of the interval ends (padded).
// If the largest return in that window (from the tile pyramid) is below d - R, `ri >= range` is
// false for every voxel of the tracklet through this pair: the pair is dropped from the work list.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool
make_pair(int64_t f, int c, int q, int L, const TrkGrid &g, const occb200_pose_t *__restrict__ poses,
          const int32_t *__restrict__ frame_sf, const occb200_sensor_t *__restrict__ sensors,
          const TabCoef *__restrict__ tabcoef, const LutCell *__restrict__ lut_pool, double vs,
          const int64_t *__restrict__ pyr_off, const float *__restrict__ pyr, int cull_on, PairHot &p) {
  const occb200_pose_t &ps = poses[f];
  const int64_t se = (int64_t)frame_sf[f] * L + c;
  const occb200_sensor_t &sn = sensors[se];
  const double rc = (double)ps.cos_p, rs = (double)ps.sin_p;
  const double Rm[9] = {rc, rs, 0, -rs, rc, 0, 0, 0, 1};
  double V[12];
  for (int k = 0; k < 12; ++k) V[k] = (double)sn.v2l[k];
  double VR[9];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) VR[3 * r + k] = V[4 * r] * Rm[k] + V[4 * r + 1] * Rm[3 + k] + V[4 * r + 2] * Rm[6 + k];
  const double c0[3] = {(double)g.mb[0] + vs / 2, (double)g.mb[1] + vs / 2, (double)g.mb[2] + vs / 2};
  const double o[3] = {(double)ps.box[0], (double)ps.box[1], (double)ps.box[2]};
  double half[3], pcen[3], A64[9], R2 = 0.0;
  for (int k = 0; k < 3; ++k) half[k] = 0.5 * (double)(g.dims[k] > 0 ? g.dims[k] - 1 : 0);
  for (int r = 0; r < 3; ++r) {
    const double b = VR[3 * r] * c0[0] + VR[3 * r + 1] * c0[1] + VR[3 * r + 2] * c0[2] + V[4 * r] * o[0] +
                     V[4 * r + 1] * o[1] + V[4 * r + 2] * o[2] + V[4 * r + 3];
    pcen[r] = b;
    for (int k = 0; k < 3; ++k) {
      A64[3 * r + k] = vs * VR[3 * r + k];
      pcen[r] += half[k] * A64[3 * r + k];
    }
  }
  for (int k = 0; k < 3; ++k) {                    // half diagonal of the centre lattice in the sensor frame
    double col2 = 0.0;
    for (int r = 0; r < 3; ++r) col2 += A64[3 * r + k] * A64[3 * r + k];
    R2 += col2 * half[k] * half[k];
  }
  // the columns of VR are orthogonal only up to the f32 inverse: 1.001 covers it, +1 mm absolute
  const double R = sqrt(R2) * 1.001 + 1e-3;
  const double rho_c = sqrt(pcen[0] * pcen[0] + pcen[1] * pcen[1]);
  const double d_c = sqrt(rho_c * rho_c + pcen[2] * pcen[2]);
  const bool has_dir = rho_c > 1e-9;
  const double ct = has_dir ? pcen[0] / rho_c : 1.0, st = has_dir ? pcen[1] / rho_c : 0.0;
  const double theta = has_dir ? atan2(pcen[1], pcen[0]) : 0.0;
  double A2[9], bc[3], M[3];
  for (int k = 0; k < 3; ++k) {
    A2[k] = ct * A64[k] + st * A64[3 + k];
    A2[3 + k] = -st * A64[k] + ct * A64[3 + k];
    A2[6 + k] = A64[6 + k];
  }
  bc[0] = ct * pcen[0] + st * pcen[1];
  bc[1] = -st * pcen[0] + ct * pcen[1];
  bc[2] = pcen[2];
  for (int r = 0; r < 3; ++r) {
    M[r] = fabs(bc[r]);
    for (int k = 0; k < 3; ++k) {
      M[r] += fabs(A2[3 * r + k]) * half[k];
      p.A[3 * r + k] = (float)A2[3 * r + k];
    }
    p.bc[r] = (float)bc[r];
  }
  const double k6 = 6.0 * 5.9604644775390625e-08;
  const double eps_y = k6 * M[1], eps_xz = k6 * fmax(M[0], M[2]);
  const bool narrow = rho_c > 2.1 * R;
  const double tmax = narrow ? fmin(R / (rho_c - R), 1.0) : 1.0;
  const int W = sn.W, H = sn.H;
  const double kcol = (double)W / 6.28318530717958647692;
  // colf = (W - 0.5) - (az + pi) / (2 pi) * W with az = theta + phi + azc  (:176-191); taken modulo W
  const double C0 = ((double)W - 0.5) - (theta + (double)sn.azc + 3.14159265358979323846) / 6.28318530717958647692 * (double)W;
  const double C0m = C0 - floor(C0 / (double)W) * (double)W;       // in [0, W)
  int cint = (int)floor(C0m);
  cint = max(0, min(cint, W - 1));
  const double phimax = narrow ? atan(tmax) : 3.14159265358979323846;
  // colf_rel = fma(phi, -kcol, c0f): one rounding at magnitude <= kcol * phimax + 1, the rounding of kcol and of
  // c0f; the reference wraps with float32(2 pi), which moves its colf by W * 3e-8 relative to an exact wrap
  const double c_col = 3.0 * 5.9604644775390625e-08 * (kcol * phimax + 2.0) + (double)W * 4e-8;
  const double katan = narrow ? (double)kAtanNarrow : (double)kAtanWide;
  const TabCoef tc = tabcoef[sn.incl_off];
  const double e15z = 1.5 * fmax(eps_xz, eps_y);
  const double d_min = d_c - R;
  bool ok = tc.ok && tc.H == H && sn.incl_mono == -1 && isfinite(eps_xz) && isfinite(eps_y) && se < (1ll << 31) &&
            W >= 16 && W < (1 << 22) && d_min > 0.05;          // (W >= 16: the narrow column wrap of fast_test)
  // the row test needs its margin below a fifth of a lookup cell (k_table_setup / fast_test)
  if (ok && !(e15z / d_min + 1.5e-6 < 0.2 * (double)tc.w)) ok = false;
  p.eps = ok ? (float)fmax(eps_xz, eps_y) : -1.f;
  p.q = q;
  p.inv_w = tc.inv_w;
  p.cell0m = tc.cell0m;
  p.ncm1 = (uint32_t)max(tc.ncell - 1, 0);
  p.last = (uint32_t)max(H - 1, 0);
  p.e15z = (float)e15z;
  p.nkcol = (float)(-kcol);
  p.c0f = (float)(C0m - (double)cint);
  p.cint = (narrow && cint < W / 2) ? cint + W : cint;         // narrow: fast_test wraps with one unsigned minimum
  p.W = W;
  p.ecolk = (float)(1.5 * (eps_y + tmax * eps_xz) * kcol);
  p.ecol = (float)((katan + 3.0e-7) * kcol + c_col);
  p.c1 = (float)(1.6 * eps_xz + 1.1 * eps_y);
  p.lut_off = (int32_t)(sn.incl_off * kLutPerRow);
  p.wide = narrow ? 0 : 1;
  p.ri_off = sn.ri_off;
  p.sens = (int32_t)se;
  p.pad = 0;
  // ---- cull test (conservative; any doubt keeps the pair)
  if (cull_on && g.status == OCCB200_OK && sn.incl_mono == -1 && H >= 1 && W >= 1) {
    // f32 geometry: the window is padded by 1e-4 rad (>> f32 error, << a pixel) and extra rows / columns
    const float Rf = (float)R;
    const float rho = (float)rho_c, d = (float)d_c;
    if (d > 1.25f * Rf) {
      const float rmin = d - Rf - 1e-3f * (1.f + d * 1e-3f);
      const float delta = asinf(Rf / d) + 1e-4f;
      const float inc_c = atan2f((float)pcen[2], rho);
      int r0 = 0, r1 = H - 1;
      if (tc.ok && tc.H == H) {
        const LutCell *lut = lut_pool + sn.incl_off * kLutPerRow;
        float sh, ch, sl, cl;
        sincosf(fminf(inc_c + delta, 1.5707f), &sh, &ch);
        sincosf(fmaxf(inc_c - delta, -1.5707f), &sl, &cl);
        const float u_hi = sh / (fabsf(sh) + ch) + 1e-5f;
        const float u_lo = sl / (fabsf(sl) + cl) - 1e-5f;
        r0 = max(row_of_u(lut, tc.inv_w, tc.cell0m, tc.ncell, u_hi) - 1, 0);
        r1 = min(row_of_u(lut, tc.inv_w, tc.cell0m, tc.ncell, u_lo) + 1, H - 1);
      }
      const int ntc = (W + kTileC - 1) / kTileC;
      long long c_lo = 0, c_hi = W - 1;            // column window, possibly beyond [0, W): taken modulo W
      if (rho > 1.05f * Rf) {
        const float daz = asinf(Rf / rho) + 1e-4f;
        const float kc = (float)kcol, cc = (float)C0m;
        const float cf_lo = cc - daz * kc, cf_hi = cc + daz * kc;
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
          for (int tcx = s0; tcx <= s1 && m < rmin; tcx += 4) {
            const float v0 = prow[tcx], v1 = prow[min(tcx + 1, s1)], v2 = prow[min(tcx + 2, s1)],
                        v3 = prow[min(tcx + 3, s1)];
            m = fmaxf(fmaxf(m, fmaxf(v0, v1)), fmaxf(v2, v3));
          }
        }
      }
      if (m < rmin) return false;
    }
  }
  return true;
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

__device__ __forceinline__ int bricks_of(int n) { return (n + kBrick - 1) / kBrick; }

// One CTA per tracklet, one thread per (frame, LiDAR) pair: the pair's record and cull decision (make_pair); the
// surviving pairs are written to the front of the tracklet's slot -- frames in strided order and, inside a frame,
// LiDARs in reference order, so that the first slice of pairs looks at the object from spread-out viewpoints;
// thread 0 fixes the final status and writes the tracklet's hot record; all threads then emit the work items:
// one per (slice of kPairsPerItem pairs, brick), appended to the slice's own list (the kernels walk the lists
// slice by slice, so the pairs that free most voxels are tested first, on every brick of the batch).
struct SetupArgs {               // what k_pair_build's set-up prologue needs (see tracklet_setup_warp)
  const int64_t *frame_pt_off;
  const int32_t *frame_kept;
  const int64_t *label_off;
  float vsf, inv_vs;
  uint32_t *bits;
  int32_t *redo_list;
  unsigned long long *redo_count;
  int32_t *dims_out;
  float *sizes_out;
  int64_t *n_unknown;
  int64_t *n_steps;
};

__global__ void __launch_bounds__(256)
k_pair_build(int T, int L, const SetupArgs su, const int64_t *__restrict__ trk_frame_off,
             const int64_t *__restrict__ brick_off, TrkGrid *__restrict__ grids, int32_t *__restrict__ trk_flags,
             const occb200_pose_t *__restrict__ poses, const int32_t *__restrict__ frame_sf,
             const occb200_sensor_t *__restrict__ sensors, const TabCoef *__restrict__ tabcoef,
             const LutCell *__restrict__ lut_pool, double vs, const int64_t *__restrict__ pyr_off,
             const float *__restrict__ pyr, int cull_on, PairHot *__restrict__ pairs_c, TrkHot *__restrict__ hot,
             int2 *__restrict__ item_map, long long bricks_total, int n_slices,
             unsigned long long *__restrict__ counter, int32_t *__restrict__ status_out) {
  __shared__ long long s_i0[kMaxSlices];
  __shared__ int s_cnt[8];
  __shared__ TrkGrid s_g;
  const int t = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {                                  // the tracklet's set-up (true grid, redo list, first status)
    const TrkGrid g0 = tracklet_setup_warp(t, lane, trk_frame_off, poses, su.frame_pt_off, su.frame_kept, su.label_off,
                                           su.vsf, su.inv_vs, 32, grids, su.bits, trk_flags, su.redo_list, su.redo_count,
                                           su.dims_out, su.sizes_out, status_out, su.n_unknown, su.n_steps);
    if (lane == 0) s_g = g0;
  }
  __syncthreads();
  const TrkGrid g = s_g;
  // the flags of a tracklet whose grid was right the first time are final; those of a corrected one are still
  // being rewritten by the redo pass (side stream) -- it is treated as OK here and k_labels folds its flags in
  int status = g.status;
  if (status == OCCB200_OK && !g.redo) {
    const int fl = g.flags;
    if (fl & 2) status = OCCB200_INDEX_ERROR;
    else if (!(fl & 1)) status = OCCB200_EMPTY_AFTER_FILTER;
  }
  const int nbx = bricks_of(g.dims[0]), nby = bricks_of(g.dims[1]), nbz = bricks_of(g.dims[2]);
  const long long nbricks = (long long)nbx * nby * nbz;
  const int64_t base = trk_frame_off[t] * L;
  const int B = (int)(trk_frame_off[t + 1] - trk_frame_off[t]);
  if (status == OCCB200_OK && (nbricks > brick_off[t + 1] - brick_off[t] || (long long)B * L > (long long)n_slices * kPairsPerItem))
    status = OCCB200_WORK_OVERFLOW;                 // a host bound (brick_off / max_pairs) is violated: never silently
  const int n = (status == OCCB200_OK) ? B * L : 0;
  int count = 0;                                    // surviving pairs so far (same value in every thread)
  for (int j0 = 0; j0 < n; j0 += 256) {             // 256 pairs per pass, one per thread
    const int j = j0 + threadIdx.x;
    PairHot p;
    bool keep = false;
    if (j < n) {
      const int pf = j / L;
      const int q = strided_frame(pf, B) * L + (j - pf * L);
      const int i = q / L;
      keep = make_pair(trk_frame_off[t] + i, q - i * L, q, L, g, poses, frame_sf, sensors, tabcoef, lut_pool, vs,
                       pyr_off, pyr, cull_on, p);
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
      const float4 *src = reinterpret_cast<const float4 *>(&p);
      float4 *d4 = reinterpret_cast<float4 *>(pairs_c + base + dst);
#pragma unroll
      for (int k = 0; k < 8; ++k) d4[k] = src[k];
    }
    count += total;
    __syncthreads();                                // s_cnt is reused by the next pass
  }
  const int nslice = (status == OCCB200_OK) ? (count + kPairsPerItem - 1) / kPairsPerItem : 0;
  if (threadIdx.x == 0) {
    TrkHot h;
    h.V = (int32_t)g.V; h.dX = g.dims[0]; h.dY = g.dims[1]; h.dZ = g.dims[2];
    h.status = status; h.nact = count;
    h.bits_off = g.bits_off; h.pairs_base = base; h.brick_base = brick_off[t];
    for (int k = 0; k < 3; ++k) h.cen[k] = 0.5f * (float)(g.dims[k] > 0 ? g.dims[k] - 1 : 0);
    h.pad = 0;
    hot[t] = h;
    status_out[t] = status;
  }
  for (int s = threadIdx.x; s < nslice; s += blockDim.x)
    s_i0[s] = (long long)atomicAdd(counter + 8 + 2 * s, (unsigned long long)nbricks);
  __syncthreads();
  const int nyz = nby * nbz;
  for (long long i = threadIdx.x; i < nbricks * nslice; i += blockDim.x) {
    const int s = (int)(i / nbricks);
    const int b = (int)(i - (long long)s * nbricks);
    const int bx = b / nyz, rem = b - bx * nyz;
    const int by = rem / nbz, bz = rem - by * nbz;
    item_map[(long long)s * bricks_total + s_i0[s] + b] = make_int2(t, bx | (by << 10) | (bz << 20));
  }
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

// atan(a) = a * P(a^2), P of degree 6: |atan(a) - a P(a^2)| <= 2.5e-7 for |a| <= 1 in exact arithmetic (checked on
// 2e6 points); Horner rounding <= 4e-7; the caller's approximate division (2 ulp of a) <= 2.4e-7.
__device__ __forceinline__ float atan_poly(float a) {
  const float s = a * a;
  float p = 0.006811790633946657f;
  p = fmaf(p, s, -0.0336042158305645f);
  p = fmaf(p, s, 0.07962366938591003f);
  p = fmaf(p, s, -0.1323334276676178f);
  p = fmaf(p, s, 0.19807815551757812f);
  p = fmaf(p, s, -0.3331736922264099f);
  p = fmaf(p, s, 0.9999961256980896f);
  return p * a;
}

// Narrow pairs: x > 0 and |y / x| <= 1 by construction (make_pair): atan2(y, x) = atan(y / x), no quadrant logic.
// Total error < 9e-7 rad; kAtanNarrow = 1e-6 is the bound used by the margins.
__device__ __forceinline__ float atan_narrow(float y, float x) { return atan_poly(y * rcp_approx(x)); }

// Full-quadrant atan2 for the wide pairs: a = min/max in [0,1], plus the quadrant fix-ups (2 roundings at <= pi:
// 2.4e-7 each).  Total < 1.4e-6 rad; kAtanWide = 2e-6 (occb200_selftest_atan2 measures both on the device).
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float r = atan_poly(mn * rcp_approx(mx));
  if (ay > ax) r = 1.57079632679489661923f - r;
  if (x < 0.f) r = 3.14159265358979323846f - r;
  return (y < 0.f) ? -r : r;
}

// ---------------------------------------------------------------------------------------------
// Brick-level cull (docs/ROUND2_BRICK_CULL.md).  Brick = the voxel centres idx in [4bx, 4bx+3] x [4by, 4by+3] x
// [4bz, 4bz+3].  Through a pair they map to a convex body with corners v_0..v_7 (corners of the FULL brick: a
// superset of a clipped one).  With c the image of the brick centre and u = c/|c|:
//   range    r_lo = min_i v_i.u  <=  |p|  <=  max_i |v_i| = r_hi   (a linear function attains its minimum, a convex
//            one its maximum, over a convex body at a corner)
//   columns  azimuth extremes of a convex body clear of the sensor's z axis are attained at corners
//   rows     z is linear, and sin(inc) = z / |p| is bracketed with r_lo / r_hi by the sign of z
// If the largest return over that pixel footprint (fine pyramid level, footprint snapped outwards to whole tiles) is
// below r_lo, `ri >= range` is false for every centre of the brick through the pair: its mask bit is set and the
// visibility kernel never evaluates the pair for the brick.  f32 geometry with generous padding (1e-4 rad, 1 mm +
// the pair's error bound, extra rows / columns); any doubt keeps the pair.  The cull only removes tests that must
// fail, so labels cannot change (flag bit 4 switches it off for the A/B parity tests).
// One thread per (work item of slice s, pair of the slice): blockIdx.y = slice.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u_of_sin(float s) {
  const float c2 = fmaxf(1.f - s * s, 1e-12f);
  return s * rcp_approx(fabsf(s) + c2 * rsqrt_approx(c2));
}

// All quantities below are conservative bounds with generous padding, so the approximate reciprocal / square root
// / arctangent (<= 2 ulp, <= 2e-6 rad) are used throughout: their error is orders of magnitude below the padding.
#ifndef OCC_CULL_MINB
#define OCC_CULL_MINB 6
#endif
#ifndef OCC_CULL_UNROLL
#define OCC_CULL_UNROLL 2
#endif
constexpr int kCullUnroll = OCC_CULL_UNROLL;
__global__ void __launch_bounds__(256, OCC_CULL_MINB)
k_brick_cull(int s_first, const int2 *__restrict__ item_map, long long bricks_total, const unsigned long long *__restrict__ counter,
             const TrkHot *__restrict__ hot, const PairHot *__restrict__ pairs, const LutCell *__restrict__ lut_pool,
             const int64_t *__restrict__ pyr_off, const float *__restrict__ pyr2, int mask_words,
             uint32_t *__restrict__ pair_mask) {
  const int s = s_first + blockIdx.y;
  const long long n_items = (long long)counter[8 + 2 * s];
  // thread -> (brick item, pair): a warp holds 32 CONSECUTIVE bricks (one tracklet, as a rule) and ONE pair, so the
  // pair record is read at a warp-uniform address (one L1 wavefront per load instead of one per distinct record:
  // with 16 pairs across the lanes the record loads alone kept the load / store unit busy for most of the kernel)
  const long long n_groups = (n_items + 31) / 32;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < n_groups * 32 * kPairsPerItem;
       gidx += (long long)gridDim.x * blockDim.x) {
    const long long grp = gidx / (32 * kPairsPerItem);
    const int within = (int)(gidx - grp * (32 * kPairsPerItem));
    const long long item = grp * 32 + (within & 31);
    if (item >= n_items) continue;
    const int k = s * kPairsPerItem + (within >> 5);
    const int2 m = __ldg(item_map + (long long)s * bricks_total + item);
    const TrkHot &h = hot[m.x];
    if (k >= h.nact) continue;
    const PairHot &p = pairs[h.pairs_base + k];
    if (!(p.eps >= 0.f)) continue;
    const int bx = m.y & 1023, by = (m.y >> 10) & 1023, bz = (m.y >> 20) & 1023;
    const int W = p.W, H = (int)p.last + 1;
    const float hs = 0.5f * (kBrick - 1);
    // image of the brick centre (rotated frame) and the three half-edge vectors a, b, c of the brick's centre
    // lattice: the 8 corners are pc +- a +- b +- c, so every corner extreme below is a sum of absolute values
    const float cx = (float)(kBrick * bx) + hs - h.cen[0], cy = (float)(kBrick * by) + hs - h.cen[1],
                cz = (float)(kBrick * bz) + hs - h.cen[2];
    const float pcx = fmaf(cz, p.A[2], fmaf(cy, p.A[1], fmaf(cx, p.A[0], p.bc[0])));
    const float pcy = fmaf(cz, p.A[5], fmaf(cy, p.A[4], fmaf(cx, p.A[3], p.bc[1])));
    const float pcz = fmaf(cz, p.A[8], fmaf(cy, p.A[7], fmaf(cx, p.A[6], p.bc[2])));
    const float ax = hs * p.A[0], ay = hs * p.A[3], az = hs * p.A[6];
    const float bx_ = hs * p.A[1], by_ = hs * p.A[4], bz_ = hs * p.A[7];
    const float cx_ = hs * p.A[2], cy_ = hs * p.A[5], cz_ = hs * p.A[8];
    const float R2 = fmaf(cz_, cz_, fmaf(cy_, cy_, fmaf(cx_, cx_, fmaf(bz_, bz_, fmaf(by_, by_, fmaf(bx_, bx_,
                     fmaf(az, az, fmaf(ay, ay, ax * ax))))))));
    const float R = R2 * rsqrt_approx(R2) * 1.001f + 1e-3f;       // half diagonal (the edge vectors are orthogonal)
    const float rho2 = fmaf(pcy, pcy, pcx * pcx), d2 = fmaf(pcz, pcz, rho2);
    const float inv_d = rsqrt_approx(d2);
    const float rho_c = rho2 * rsqrt_approx(rho2), d = d2 * inv_d;
    if (!(d > 1.25f * R && rho_c > 1.05f * R)) continue;
    const float ux = pcx * inv_d, uy = pcy * inv_d, uz = pcz * inv_d;
    const float slack = 1e-3f + 2.f * p.eps + 1e-5f * d;
    // range: min over corners of v.u = pc.u - sum |e.u|;  max over corners of |v| <= |pc| + half diagonal
    const float r_lo = d - (fabsf(fmaf(az, uz, fmaf(ay, uy, ax * ux))) + fabsf(fmaf(bz_, uz, fmaf(by_, uy, bx_ * ux))) +
                            fabsf(fmaf(cz_, uz, fmaf(cy_, uy, cx_ * ux)))) - slack;
    const float r_hi = d + R + slack;
    const float zext = fabsf(az) + fabsf(bz_) + fabsf(cz_) + slack;
    const float zmin = pcz - zext, zmax = pcz + zext;
    // azimuth relative to the brick centre's: tan = cross(pc, v) / dot(pc, v) in the xy plane; over the corners
    // |cross| <= sum |cross(pc, e)| and dot >= rho^2 - sum |dot(pc, e)| (> 0 because rho_c > 1.05 R)
    const float crs = fabsf(fmaf(pcx, ay, -(pcy * ax))) + fabsf(fmaf(pcx, by_, -(pcy * bx_))) + fabsf(fmaf(pcx, cy_, -(pcy * cx_)));
    const float dmin = rho2 - (fabsf(fmaf(pcy, ay, pcx * ax)) + fabsf(fmaf(pcy, by_, pcx * bx_)) + fabsf(fmaf(pcy, cy_, pcx * cx_)));
    if (!(r_lo > 0.f) || !(dmin > 0.125f * rho2)) continue;
    const float tq = crs * rcp_approx(dmin);                      // <= 8 R / rho: finite
    // rows: sin(inc) = z / |p| bracketed by the corner extremes, through the u-space lookup like the pair cull;
    // row_of_u counts the boundaries above u exactly, and u is padded by 2e-4 (>> the f32 evaluation error)
    const float inv_lo = rcp_approx(r_lo), inv_hi = rcp_approx(r_hi);
    const float s_hi = fminf(fmaxf(zmax * (zmax > 0.f ? inv_lo : inv_hi), -1.f), 1.f);
    const float s_lo = fminf(fmaxf(zmin * (zmin > 0.f ? inv_hi : inv_lo), -1.f), 1.f);
    const float u_hi = u_of_sin(s_hi) + 2e-4f, u_lo = u_of_sin(s_lo) - 2e-4f;
    const LutCell *lut = lut_pool + p.lut_off;
    const int ncell = (int)p.ncm1 + 1;
    const int r0 = max(min(row_of_u(lut, p.inv_w, p.cell0m, ncell, u_hi), H - 1), 0);
    const int r1 = max(min(row_of_u(lut, p.inv_w, p.cell0m, ncell, u_lo), H - 1), 0);
    // columns: colf_rel = c0f - kcol * phi with phi = atan2(pcy, pcx) +- (atan(tq) + 1e-4)
    const float kc = -p.nkcol;
    const float phi_c = p.wide ? atan2_fast(pcy, pcx) : atan_narrow(pcy, pcx);
    const float dphi = atan_poly(fminf(tq, 1.f)) + (tq > 1.f ? 1.f : 0.f) + 1e-4f;   // tq > 1: any bound >= atan(tq) keeps it safe
    const float cf_lo = p.c0f - (phi_c + dphi) * kc;
    const float cf_hi = p.c0f - (phi_c - dphi) * kc;
    const int q_lo = p.cint + (int)floorf(cf_lo) - 1, q_hi = p.cint + (int)ceilf(cf_hi) + 1;
    const int len = q_hi - q_lo + 1;
    if (!(len > 0 && len < W / 2)) continue;
    // fine tiles covering rows [r0, r1] and columns [q_lo, q_hi] modulo W: segment 1 = tiles [ta, tb], segment 2
    // (wrapped past the seam) = tiles [0, tw]
    const int pitch2 = 4 * ((W + kTileC - 1) / kTileC);
    const float4 *pimg = reinterpret_cast<const float4 *>(pyr2 + 16 * pyr_off[p.sens]);
    if (!(q_lo > -W && q_lo < 3 * W)) continue;                   // (cannot happen for finite inputs: keep the pair)
    int a0 = q_lo + ((q_lo < 0) ? W : 0);                         // q_lo modulo W without the integer division
    a0 -= (a0 >= 2 * W) ? 2 * W : ((a0 >= W) ? W : 0);
    const int ta = a0 / kFineC, tb = (min(a0 + len, W) - 1) / kFineC;
    const int tw = (a0 + len > W) ? (a0 + len - W - 1) / kFineC : -1;
    // every tile of the footprint, four tiles per 16-byte load, no early exit (the loads are independent of the
    // running maximum).  Columns outermost: which of a load's four tiles belong to the footprint depends on the
    // column group only, so the row loop is load + 4 max and the mask is applied once per group.
    const int tr0 = r0 / kFineR, tr1 = r1 / kFineR;
    const int rstride = pitch2 >> 2;                              // float4 per tile row
    float mx = 0.f;
    auto scan = [&](int g0, int g1, int lo, int hi) {             // column groups [g0, g1], tiles [lo, hi] count
      for (int g = g0; g <= g1; ++g) {
        const float4 *pcol = pimg + (tr0 * rstride + g);
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll kCullUnroll
        for (int tr = tr0; tr <= tr1; ++tr, pcol += rstride) {
          const float4 v = __ldg(pcol);
          m0 = fmaxf(m0, v.x);
          m1 = fmaxf(m1, v.y);
          m2 = fmaxf(m2, v.z);
          m3 = fmaxf(m3, v.w);
        }
        const int c = 4 * g;
        mx = fmaxf(mx, (c >= lo && c <= hi) ? m0 : 0.f);
        mx = fmaxf(mx, (c + 1 >= lo && c + 1 <= hi) ? m1 : 0.f);
        mx = fmaxf(mx, (c + 2 >= lo && c + 2 <= hi) ? m2 : 0.f);
        mx = fmaxf(mx, (c + 3 >= lo && c + 3 <= hi) ? m3 : 0.f);
      }
    };
    scan(ta >> 2, tb >> 2, ta, tb);
    if (tw >= 0) scan(0, tw >> 2, 0, tw);
    if (mx < r_lo) {
      const int lb = (bx * bricks_of(h.dY) + by) * bricks_of(h.dZ) + bz;
      atomicOr(pair_mask + (h.brick_base + lb) * mask_words + (k >> 5), 1u << (k & 31));
    }
  }
}

// float -> nearest integer (ties to even) without the conversion unit: valid for |x| < 2^22; anything else
// (NaN included) gives an arbitrary integer, which every caller clamps before using it as an index.
constexpr float kMagic = 12582912.f;              // 1.5 * 2^23
__device__ __forceinline__ int magic_int(float biased) { return __float_as_int(biased) - 0x4B400000; }

// One fast test of the voxel centre at lattice offset (dx, dy, dz) = idx - cen.
// Returns 2 = certainly free, 0 = certainly not free, 1 = undecided (recheck in f64).
//
//   position  p' = bc + A d (3 FMAs per component): |p' - p_ref| <= eps_y in y', eps_xz in x' and z' (make_pair)
//   row       u = z' / (|z'| + rho); |u - u_ref| <= 1.42 max(eps) / r + evaluation (~6 ulp of 1).  The lookup cell
//             (nearest) holds the only boundary b within its extended interval and the row count h above it:
//             row = h + (b > u), accepted iff |u - b| > margin.  Proof of exactness: u lies within 0.5005 cell widths
//             of the cell centre, the reference's value within margin (< 0.2 widths, checked per pair) of u, so
//             both sit inside the extended interval (+-0.75 widths), where b is the only boundary.
//   column    phi = atan(y'/x') (narrow) or atan2 (wide); colf_rel = c0f - kcol phi; col = cint + rint(colf_rel)
//             modulo W, accepted iff colf_rel is farther from a half-integer than ecolk / rho + ecol
//   range     free iff ri >= |p_ref|;  | r - |p_ref| | <= 1.42 eps_xz + eps_y + 6.5e-7 r
template <bool WIDE>
__device__ __forceinline__ int fast_test(const PairHot &p, float dx, float dy, float dz,
                                         const int2 *__restrict__ lut, const float *__restrict__ ri_img) {
  const float px = fmaf(dz, p.A[2], fmaf(dy, p.A[1], fmaf(dx, p.A[0], p.bc[0])));
  const float py = fmaf(dz, p.A[5], fmaf(dy, p.A[4], fmaf(dx, p.A[3], p.bc[1])));
  const float pz = fmaf(dz, p.A[8], fmaf(dy, p.A[7], fmaf(dx, p.A[6], p.bc[2])));
  const float s2 = fmaf(py, py, px * px);
  const float r2 = fmaf(pz, pz, s2);
  const float inv_rho = rsqrt_approx(s2);
  const float inv_r = rsqrt_approx(r2);

  // ---- row
  const float u = pz * rcp_approx(fmaf(s2, inv_rho, fabsf(pz)));
  const unsigned cell = min((unsigned)magic_int(fmaf(u, p.inv_w, p.cell0m) + kMagic), p.ncm1);
  const int2 lc = __ldg(lut + cell);                          // (boundary, rows above it)
  const float b = __int_as_float(lc.x);
  const unsigned row = (unsigned)lc.y + ((b > u) ? 1u : 0u);
  const bool ok_row = fabsf(u - b) > fmaf(p.e15z, inv_r, 1.5e-6f);

  // ---- column
  const float phi = WIDE ? atan2_fast(py, px) : atan_narrow(py, px);
  const float colf = fmaf(phi, p.nkcol, p.c0f);
  const float cb = colf + kMagic;                             // |colf| <= W / 2 + 1 < 2^22
  const float cr = cb - kMagic;                               // == rintf(colf)
  const bool ok_col = fabsf(colf - cr) + fmaf(p.ecolk, inv_rho, p.ecol) < 0.5f;
  unsigned col;
  if (WIDE) {
    int c = magic_int(cb) + p.cint;                           // cint in [0, W), |colf| <= W / 2 + 1
    c += (c < 0) ? p.W : 0;                                   // fmod(round(colf), W) (:191) and
    c -= (c >= p.W) ? p.W : 0;                                // negative index wrap (:543)
    col = (unsigned)c;
  } else {
    // narrow pairs: |colf| <= W / 8 + 1 and make_pair stores cint + W where cint < W / 2, so the sum lies in
    // (0, 2 W): ONE unsigned minimum does the wrap (col - W wraps around to a huge number when col < W)
    col = (unsigned)(magic_int(cb) + p.cint);
    col = min(col, col - (unsigned)p.W);
  }
  col = min(col, (unsigned)(p.W - 1));                        // (only a NaN position gets here out of range)

  // ---- range
  const float ri = __ldg(ri_img + (min(row, p.last) * (unsigned)p.W + col));
  const float r = r2 * inv_r;
  const float dd = ri - r;
  const bool ok_rng = fabsf(dd) > fmaf(r, 6.5e-7f, p.c1);     // ri == 0 (no return): dd = -r, certainly not free
  return (ok_row && ok_col && ok_rng) ? ((dd > 0.f) ? 2 : 0) : 1;
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

struct VisArgs {                 // what the visibility kernel needs (passed by value: one constant-bank block)
  int L;
  int mask_words;
  int s_lo, s_hi;                // slices of pairs this launch walks
  int ticket;                    // which ticket counter it draws from (counter[3 + ticket])
  int pad;
  long long bricks_total;
  long long queue_cap;
  double vs;
  const int64_t *trk_frame_off;
  const occb200_pose_t *poses;
  const int32_t *frame_sf;
  const occb200_sensor_t *sensors;
  const float *incl_pool;
  const float *ri_pool;
  const TrkGrid *grids;
  unsigned long long *counter;
  const uint32_t *bits;
  uint32_t *free_brick;
  const uint32_t *unk_brick;
  const uint32_t *pair_mask;
  const int2 *item_map;
  const TrkHot *hot;
  const PairHot *pairs;
  const LutCell *lut_pool;
  int4 *queue;
  int64_t *n_steps;
};

// The pair loop of one work item: VPL voxels per lane (lattice offsets d*, brick-local ids vj, -1 = none).
// Returns the bits of the voxels proven free (bit v = this lane's voxel v).
template <int VPL, int PP, bool MIXED>
__device__ __forceinline__ unsigned run_pairs(const VisArgs &a, int t, const TrkHot &h, int bx, int by, int bz, int lb,
                                              unsigned live, int k0, unsigned todo, const float (&dx)[VPL],
                                              const float (&dy)[VPL], const float (&dz)[VPL], const int (&vj)[VPL],
                                              unsigned &steps, const PairHot *__restrict__ staged, unsigned &iters) {
  // PP pairs per loop iteration: their VPL x PP tests are independent dependent-load chains (record -> lookup cell
  // -> pixel) that overlap; a voxel freed by the first pair of an iteration still pays the second one's test.
  const int lane = threadIdx.x & 31;
  unsigned found = 0u;
  while (live) {
    if (!__any_sync(0xffffffffu, todo != 0u)) break;           // every voxel of the item is settled
    ++iters;
    int res[PP][VPL];
    const PairHot *pcs[PP];
    bool has[PP];
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      has[p] = live != 0u;
      const int kk = has[p] ? __ffs(live) - 1 : 0;
      live &= live - 1u;                                         // (0 stays 0)
      pcs[p] = staged + kk;
    }
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      // (an absent second pair repeats record 0: straight-line code, its results are dropped below)
      const PairHot &pc = *pcs[p];                               // the warp's own shared-memory copy (broadcast reads)
      const int2 *lut = reinterpret_cast<const int2 *>(a.lut_pool) + pc.lut_off;
      const float *ri_img = a.ri_pool + pc.ri_off;
      // materialise the bases as 64-bit registers: per-test addresses are then ONE imad.wide each
      asm volatile("" : "+l"(lut), "+l"(ri_img));
      // all VPL tests are evaluated unconditionally (no divergence, their dependent chains interleave);
      // results of voxels this lane does not need are discarded
      if (MIXED && pc.wide) {                                      // (a branch here keeps the pairs' chains apart)
#pragma unroll
        for (int v = 0; v < VPL; ++v) res[p][v] = fast_test<true>(pc, dx[v], dy[v], dz[v], lut, ri_img);
      } else {
#pragma unroll
        for (int v = 0; v < VPL; ++v) res[p][v] = fast_test<false>(pc, dx[v], dy[v], dz[v], lut, ri_img);
      }
    }
    unsigned und = 0u;                                           // bit p * VPL + v: that test is undecided
#pragma unroll
    for (int p = 0; p < PP; ++p) {
      if (!has[p]) continue;
      const bool fast_ok = pcs[p]->eps >= 0.f;                    // eps < 0: no fast path through this pair
      steps += __popc(todo);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {                              // (branch-free: selects only)
        const unsigned need = (todo >> v) & 1u;
        const unsigned fr = (fast_ok && res[p][v] == 2) ? need : 0u;
        const unsigned ud = (!fast_ok || res[p][v] == 1) ? need : 0u;
        found |= fr << v;
        todo &= ~(fr << v);
        und |= ud << (p * VPL + v);
      }
    }
    if (__any_sync(0xffffffffu, und != 0u)) {                  // rare: queue the undecided tests (warp-aggregated)
#pragma unroll
      for (int p = 0; p < PP; ++p) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const bool mine = (und >> (p * VPL + v)) & 1u;
          const unsigned umask = __ballot_sync(0xffffffffu, mine);
          if (!umask) continue;
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(a.counter + 1, (unsigned long long)__popc(umask));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (mine) {
            const int j = vj[v];
            const int q = pcs[p]->q;
            const int f = ((kBrick * bx + (j >> 4)) * h.dY + kBrick * by + ((j >> 2) & 3)) * h.dZ + kBrick * bz + (j & 3);
            const unsigned long long slot = base + __popc(umask & ((1u << lane) - 1u));
            if (slot < (unsigned long long)a.queue_cap) {
              a.queue[slot] = make_int4(t, f, q, lb * 64 + j);
            } else if (exact_from_ids(t, f, q, a.L, a.vs, a.grids, a.trk_frame_off, a.poses, a.frame_sf, a.sensors,
                                      a.incl_pool, a.ri_pool)) {     // queue full: decide right here
              found |= 1u << v;
              todo &= ~(1u << v);
            }
          }
        }
      }
    }
  }
  return found;
}

// Every warp is on its own: it claims a work item from the atomic ticket of the current slice, tests, and ORs the
// voxels it proved free into the global free bitset (brick order).  No shared memory, no barriers.
//   item = one 4x4x4 brick of a tracklet x one slice of kPairsPerItem of its surviving pairs, minus the pairs
//   k_brick_cull masked for the brick.  The lists are walked slice by slice: slice 0 (spread-out viewpoints)
//   frees most of the voxels that can be freed at all; an item re-reads the free bits when it starts, so later
//   slices never test a voxel an earlier one has freed.  A brick with more than 32 undecided voxels runs two per
//   lane; otherwise the undecided voxels are dealt one per lane (dense lanes).
// Labels are written afterwards by k_labels from the occupancy and free bitsets.
// One warp per brick of the slice-0 list (every brick of every tracklet that has a surviving pair): the brick's
// voxels that lie inside the grid and hold no point, as two words in brick order -- what every work item of the
// brick starts from (an item then needs two 8-byte loads instead of 64 bit extractions from the linear bitset).
// Runs after the redo crop pass (the occupancy bits are final), next to k_brick_cull.
__global__ void __launch_bounds__(256)
k_brick_unknown(const int2 *__restrict__ item_map, const unsigned long long *__restrict__ counter,
                const TrkHot *__restrict__ hot, const uint32_t *__restrict__ bits, uint32_t *__restrict__ unk_brick) {
  const long long n_items = (long long)counter[8];
  const int lane = threadIdx.x & 31;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_items;
       i += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int2 m = __ldg(item_map + i);
    const TrkHot &h = hot[m.x];
    const int bx = m.y & 1023, by = (m.y >> 10) & 1023, bz = (m.y >> 20) & 1023;
    const int lb = (bx * ((h.dY + kBrick - 1) / kBrick) + by) * ((h.dZ + kBrick - 1) / kBrick) + bz;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int j = 32 * v + lane;
      const int x = kBrick * bx + (j >> 4), y = kBrick * by + ((j >> 2) & 3), z = kBrick * bz + (j & 3);
      bool need = x < h.dX && y < h.dY && z < h.dZ;
      if (need) {
        const int f = (x * h.dY + y) * h.dZ + z;
        need = !((__ldg(bits + h.bits_off + (f >> 5)) >> (f & 31)) & 1u);
      }
      const unsigned w = __ballot_sync(0xffffffffu, need);
      if (lane == 0) unk_brick[2 * (h.brick_base + lb) + v] = w;
    }
  }
}

// 1-D bulk copy global -> shared (TMA engine) completing on an mbarrier, and the warp-wide wait for it.
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's earlier generic-proxy reads come first
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned long long *bar, unsigned phase) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(b), "r"(phase) : "memory");
}

#if OCC_VISDEBUG
__device__ long long g_visdbg[8 * 148 * 8 * 8];    // per warp: t_start, t_end, items, iterations, longest item (cycles, its iterations), smid
#endif
__global__ void __launch_bounds__(32 * kFastWarps, OCC_MINB) k_visibility(const VisArgs a) {
#if OCC_VISDEBUG
  const long long dbg_t0 = clock64();
  long long dbg_items = 0, dbg_iters = 0, dbg_max = 0, dbg_max_it = 0;
  unsigned long long dbg_g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
#endif
  __shared__ long long s_base[kMaxSlices + 1];
  __shared__ __align__(16) PairHot s_pairs[kFastWarps][kPairsPerItem];
  __shared__ unsigned char s_sel[kFastWarps][32];
#if OCC_BULK_STAGE
  __shared__ __align__(8) unsigned long long s_bar[kFastWarps];   // one mbarrier per warp: its staging copy
  if ((threadIdx.x & 31) == 0) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(&s_bar[threadIdx.x >> 5]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned bar_phase = 0;
#endif
  const int ns = a.s_hi - a.s_lo;                  // this launch walks the lists of slices [s_lo, s_hi)
  if (threadIdx.x < ns) s_base[threadIdx.x + 1] = (long long)a.counter[8 + 2 * (a.s_lo + threadIdx.x)];
  __syncthreads();
  if (threadIdx.x == 0) {
    s_base[0] = 0;
    for (int s = 0; s < ns; ++s) s_base[s + 1] += s_base[s];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long n_items = s_base[ns];
  // Work assignment: the per-slice lists are walked as one sequence (slice 0, the heavy items, first).  A warp's
  // first item is static (its global warp index: no atomic while every warp of the grid starts at once); every
  // later one is a ticket from a global counter, requested BEFORE the current item is processed so that the
  // atomic's round trip hides behind it.  (Static rounds left the slowest CTA as the kernel's tail: SMs were
  // active 70 % of the kernel's duration.)
  // The items of the LAST round (as many as there are warps) are cut into kTailSplit tickets of kPairsPerItem /
  // kTailSplit pairs each: whatever a warp starts when the tickets run out is then a quarter of an item, and the
  // kernel's drain is as long as 4 pairs instead of 16 (measured on C2: every warp busy until 24 us, the last one
  // until 51 us -- one 16-pair item of a never-freed brick).  The parts of an item run on different warps at the
  // same time; they only lose the early exit between them, and late items rarely free anything.
  const long long n_static = (long long)gridDim.x * kFastWarps;
  const long long n_split = min(n_items, n_static);
  const long long n_whole = n_items - n_split;
  const long long n_tickets = n_whole + n_split * kTailSplit;
  long long g = (long long)blockIdx.x * kFastWarps + (threadIdx.x >> 5);
  int s = 0;
  for (;;) {
    if (g >= n_tickets) break;
    unsigned long long nxt = 0;
    if (lane == 0) nxt = atomicAdd(a.counter + 3 + a.ticket, 1ull);
    long long gi = g;
    unsigned part_mask = 0xffffffffu;                          // which pairs of the slice this ticket covers
    if (g >= n_whole) {
      const long long r = g - n_whole;
      gi = n_whole + r / kTailSplit;
      constexpr int kPart = kPairsPerItem / kTailSplit;
      part_mask = ((1u << kPart) - 1u) << ((int)(r % kTailSplit) * kPart);
    }
    s = 0;
    while (gi >= s_base[s + 1]) ++s;
    const long long item = gi - s_base[s];
    s += a.s_lo;
#if OCC_VISDEBUG
    const long long dbg_ti = clock64();
    unsigned dbg_steps_lane = 0;
#endif
    const int2 *items = a.item_map + (long long)s * a.bricks_total;
    do {
      const int2 m = __ldg(items + item);
      const int t = m.x;
      const int bx = m.y & 1023, by = (m.y >> 10) & 1023, bz = (m.y >> 20) & 1023;
      const TrkHot h = load64(a.hot + t);
      const int k0 = s * kPairsPerItem;
      const int npair = min(h.nact - k0, kPairsPerItem);         // >= 1 by construction of the lists
#if OCC_BULK_STAGE
      // the slice's pair records start their way into the warp's slot now, behind the mask / bitset loads below
      __syncwarp();                                              // the previous item's reads of the slot are done
      if (lane == 0)
        bulk_copy_g2s(s_pairs[threadIdx.x >> 5], a.pairs + h.pairs_base + k0, (unsigned)npair * (unsigned)sizeof(PairHot),
                      &s_bar[threadIdx.x >> 5]);
      const unsigned my_phase = bar_phase;
      bar_phase ^= 1u;
#endif
      const int lb = (bx * ((h.dY + kBrick - 1) / kBrick) + by) * ((h.dZ + kBrick - 1) / kBrick) + bz;
      const long long gb = h.brick_base + lb;
      const unsigned mw = __ldg(a.pair_mask + gb * a.mask_words + (k0 >> 5));
      const unsigned live = ~(mw >> (k0 & 31)) & (npair >= 32 ? 0xffffffffu : ((1u << npair) - 1u)) & part_mask;
#if OCC_BULK_STAGE
      if (!live) { bulk_wait(&s_bar[threadIdx.x >> 5], my_phase); break; }
#else
      if (!live) break;
#endif
      // undecided = inside the grid, holds no point (k_brick_unknown), not yet proven free; voxel j = 32 v + lane of
      // the brick sits at (j >> 4, (j >> 2) & 3, j & 3)
      unsigned und[2];
      {
        const uint2 unk = __ldg(reinterpret_cast<const uint2 *>(a.unk_brick + 2 * gb));
        const unsigned f0 = *(volatile const uint32_t *)(a.free_brick + 2 * gb);
        const unsigned f1 = *(volatile const uint32_t *)(a.free_brick + 2 * gb + 1);
        und[0] = unk.x & ~f0;
        und[1] = unk.y & ~f1;
      }
      const int n0 = __popc(und[0]), n = n0 + __popc(und[1]);
#if OCC_BULK_STAGE
      if (n == 0) { bulk_wait(&s_bar[threadIdx.x >> 5], my_phase); break; }
#else
      if (n == 0) break;
#endif
      unsigned steps = 0, iters = 0;
      // the slice's pair records (<= kPairsPerItem x 128 bytes, contiguous) into the warp's own shared-memory slot:
      // four independent 16-byte loads per lane, all records at once, instead of one dependent 128-byte warp-uniform
      // load per loop iteration
      PairHot *staged = s_pairs[threadIdx.x >> 5];
#if OCC_BULK_STAGE
      bulk_wait(&s_bar[threadIdx.x >> 5], my_phase);
#else
      {
        const float4 *src = reinterpret_cast<const float4 *>(a.pairs + h.pairs_base + k0);
        float4 *dst = reinterpret_cast<float4 *>(staged);
        __syncwarp();                                            // the previous item's reads are done
#pragma unroll
        for (int q = 0; q < kPairsPerItem * 8 / 32; ++q)
          if (q * 32 + lane < npair * 8) dst[q * 32 + lane] = __ldg(src + q * 32 + lane);
        __syncwarp();
      }
#endif
      // pairs through which the object spans more than +-45 degrees of azimuth (an object next to the sensor) need
      // the full-quadrant arctangent: an item that holds one takes the branching loop, one pair per iteration
      const bool any_wide = __any_sync(0xffffffffu, lane < npair && staged[lane].wide != 0);
      if (n > 32) {                                              // two voxels per lane, natural mapping
        float dx[2], dy[2], dz[2];
        int vj[2];
        unsigned todo = 0u;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int j = 32 * v + lane;
          vj[v] = j;
          dx[v] = (float)(kBrick * bx + (j >> 4)) - h.cen[0];
          dy[v] = (float)(kBrick * by + ((j >> 2) & 3)) - h.cen[1];
          dz[v] = (float)(kBrick * bz + (j & 3)) - h.cen[2];
          todo |= ((und[v] >> lane) & 1u) << v;
        }
        const unsigned found = any_wide
            ? run_pairs<2, 1, true>(a, t, h, bx, by, bz, lb, live, k0, todo, dx, dy, dz, vj, steps, staged, iters)
            : run_pairs<2, OCC_PP2, false>(a, t, h, bx, by, bz, lb, live, k0, todo, dx, dy, dz, vj, steps, staged, iters);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const unsigned fw = __ballot_sync(0xffffffffu, (found >> v) & 1u);
          if (fw && lane == 0) atomicOr(a.free_brick + 2 * gb + v, fw);
        }
      } else {                                                   // one voxel per lane: lane i takes the i-th undecided
        const bool mine = lane < n;
        // every undecided voxel writes its id to the slot of its rank (the warp's 32-byte list), lane i reads slot i
        unsigned char *sel = s_sel[threadIdx.x >> 5];
        const unsigned lt = (1u << lane) - 1u;
        if ((und[0] >> lane) & 1u) sel[__popc(und[0] & lt)] = (unsigned char)lane;
        if ((und[1] >> lane) & 1u) sel[n0 + __popc(und[1] & lt)] = (unsigned char)(32 + lane);
        __syncwarp();
        float dx[1], dy[1], dz[1];
        int vj[1];
        vj[0] = mine ? (int)sel[lane] : 0;
        __syncwarp();                                            // (the next item rewrites the list)
        dx[0] = (float)(kBrick * bx + (vj[0] >> 4)) - h.cen[0];
        dy[0] = (float)(kBrick * by + ((vj[0] >> 2) & 3)) - h.cen[1];
        dz[0] = (float)(kBrick * bz + (vj[0] & 3)) - h.cen[2];
        const unsigned found = any_wide
            ? run_pairs<1, 1, true>(a, t, h, bx, by, bz, lb, live, k0, mine ? 1u : 0u, dx, dy, dz, vj, steps, staged, iters)
            : run_pairs<1, OCC_PP1, false>(a, t, h, bx, by, bz, lb, live, k0, mine ? 1u : 0u, dx, dy, dz, vj, steps, staged, iters);
        if (found) atomicOr(a.free_brick + 2 * gb + (vj[0] >> 5), 1u << (vj[0] & 31));
      }
      if (a.n_steps) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd((unsigned long long *)&a.n_steps[t], (unsigned long long)steps);
      }
#if OCC_VISDEBUG
      dbg_iters += iters;
#endif
    } while (0);
#if OCC_VISDEBUG
    {
      const long long dt = clock64() - dbg_ti;
      ++dbg_items;
      if (dt > dbg_max) { dbg_max = dt; dbg_max_it = s; }
    }
#endif
    g = n_static + (long long)__shfl_sync(0xffffffffu, nxt, 0);
  }
#if OCC_VISDEBUG
  if (lane == 0 && a.ticket == 0) {
    const long long w = (long long)blockIdx.x * kFastWarps + (threadIdx.x >> 5);
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long *d = g_visdbg + 8 * w;
    unsigned long long dbg_g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g1));
    d[0] = clock64() - dbg_t0; d[1] = (long long)dbg_g1; d[2] = dbg_items; d[3] = dbg_iters; d[4] = dbg_max; d[5] = dbg_max_it; d[6] = smid;
    d[7] = (long long)dbg_g0;
  }
#endif
}
#if OCC_VISDEBUG
extern "C" int occb200_debug_visibility(long long *out_host, int n) {
  return cudaMemcpyFromSymbol(out_host, g_visdbg, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
#endif

__global__ void __launch_bounds__(256)
k_visibility_recheck(int L, const int64_t *__restrict__ trk_frame_off, const occb200_pose_t *__restrict__ poses,
                     const int32_t *__restrict__ frame_sf, const occb200_sensor_t *__restrict__ sensors,
                     const float *__restrict__ incl_pool, const float *__restrict__ ri_pool, double vs,
                     const TrkGrid *__restrict__ grids, const TrkHot *__restrict__ hot,
                     const unsigned long long *__restrict__ counter, const int4 *__restrict__ queue,
                     long long queue_cap, uint32_t *__restrict__ free_brick, int64_t *__restrict__ n_steps) {
  const long long n = min((long long)counter[1], queue_cap);
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int4 it = queue[k];
    const int t = it.x, f = it.y, q = it.z;
    const TrkGrid &g = grids[t];
    uint32_t *word = free_brick + 2 * (hot[t].brick_base + (it.w >> 6)) + ((it.w >> 5) & 1);
    const uint32_t bit = 1u << (it.w & 31);
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

// labels from the two bitsets: 1 occupied, 2 free, 0 unknown (occ_annotate.py:558-563); grid (tracklet, 8)
__global__ void __launch_bounds__(256)
k_labels(const TrkHot *__restrict__ hot, const TrkGrid *__restrict__ grids, const int32_t *__restrict__ trk_flags,
         const int64_t *__restrict__ label_off, const uint32_t *__restrict__ bits,
         const uint32_t *__restrict__ free_brick, int32_t *__restrict__ labels, uint8_t *__restrict__ labels_u8,
         int64_t *__restrict__ n_unknown, int32_t *__restrict__ status_out) {
  const int t = blockIdx.x;
  const TrkHot h = hot[t];
  int status = h.status;
  if (status == OCCB200_OK && grids[t].redo) {      // corrected tracklet: its flags are final only now
    const int fl = trk_flags[t];
    if (fl & 2) status = OCCB200_INDEX_ERROR;
    else if (!(fl & 1)) status = OCCB200_EMPTY_AFTER_FILTER;
    if (status != OCCB200_OK && threadIdx.x == 0 && blockIdx.y == 0) status_out[t] = status;
  }
  if (status != OCCB200_OK) return;
  const int yz = h.dY * h.dZ, nby = (h.dY + kBrick - 1) / kBrick, nbz = (h.dZ + kBrick - 1) / kBrick;
  const int64_t lo = label_off[t];
  int unk = 0;
  for (int f = blockIdx.y * blockDim.x + threadIdx.x; f < h.V; f += gridDim.y * blockDim.x) {
    const bool occ = (bits[h.bits_off + (f >> 5)] >> (f & 31)) & 1u;
    const int x = f / yz, rem = f - x * yz;
    const int y = rem / h.dZ, z = rem - y * h.dZ;
    const int64_t gb = h.brick_base + ((x >> 2) * nby + (y >> 2)) * nbz + (z >> 2);
    const int j = ((x & 3) << 4) | ((y & 3) << 2) | (z & 3);
    const bool fr = (free_brick[2 * gb + (j >> 5)] >> (j & 31)) & 1u;
    const int lab = occ ? 1 : (fr ? 2 : 0);
    if (labels) labels[lo + f] = lab;
    if (labels_u8) labels_u8[lo + f] = (uint8_t)lab;
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

// self-test hook: max |atan - atan2| over n pseudo-random f32 pairs (device-side check of kAtanWide / kAtanNarrow)
__global__ void k_selftest_atan2(long long n, unsigned long long seed, int narrow, unsigned long long *__restrict__ max_err) {
  double worst = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
    h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
    const float mag_x = exp2f((float)((h >> 8) & 31) - 20.f), mag_y = exp2f((float)((h >> 13) & 31) - 20.f);
    float x = ((float)((h >> 20) & 0xfffff) / 524288.f - 1.f) * mag_x;
    float y = ((float)((h >> 40) & 0xfffff) / 524288.f - 1.f) * mag_y;
    if (narrow) {                                  // the narrow path's domain: x > 0, |y| <= x
      x = fabsf(x);
      if (fabsf(y) > x) { const float t = x; x = fabsf(y); y = (y < 0.f) ? -t : t; }
    }
    if (x == 0.f && y == 0.f) continue;
    const double got = narrow ? (double)atan_narrow(y, x) : (double)atan2_fast(y, x);
    worst = fmax(worst, fabs(got - atan2((double)y, (double)x)));
  }
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(max_err, (unsigned long long)__double_as_longlong(worst));
}

}  // namespace occb200

using namespace occb200;

extern "C" int64_t occb200_annotate_workspace_bytes(int32_t T, int64_t F, int64_t total_label_slots, int64_t SF,
                                                    int32_t L, int64_t incl_len, int64_t pyr_tiles, int64_t bricks,
                                                    int32_t max_pairs) {
  return ws_layout(T, F, total_label_slots, SF, L, incl_len, pyr_tiles, bricks, max_pairs, nullptr, nullptr);
}

extern "C" int64_t occb200_pyramid_tiles(int32_t H, int32_t W) {
  return (int64_t)((H + kTileR - 1) / kTileR) * ((W + kTileC - 1) / kTileC);
}

// HOST helper: bricks of a grid of (at most) X x Y x Z voxels.
extern "C" int64_t occb200_grid_bricks(int32_t X, int32_t Y, int32_t Z) {
  return (int64_t)((X + kBrick - 1) / kBrick) * ((Y + kBrick - 1) / kBrick) * ((Z + kBrick - 1) / kBrick);
}

extern "C" int occb200_annotate_batch(const occb200_annotate_args_t *a, int64_t total, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(a != nullptr, "args is NULL");
  OCC_REQUIRE(a->T >= 0 && a->F >= 0 && a->L >= 1, "bad T/F/L");
  OCC_REQUIRE(a->SF >= 0 && a->incl_len >= 0 && a->pyr_tiles >= 0, "bad SF / incl_len / pyr_tiles");
  OCC_REQUIRE(a->point_stride >= 3, "point_stride must be >= 3");
  OCC_REQUIRE(a->voxel_size > 0, "voxel_size must be positive");
  OCC_REQUIRE(a->labels != nullptr || a->labels_u8 != nullptr, "labels and labels_u8 are both NULL");
  if (a->T == 0) return 0;
  OCC_REQUIRE(a->F == 0 || a->frame_trk != nullptr, "frame_trk is NULL");
  OCC_REQUIRE(a->F == 0 || a->SF > 0, "tracklet-frames without sensor frames");
  OCC_REQUIRE(a->brick_off != nullptr && a->bricks >= 0, "brick_off is NULL");
  OCC_REQUIRE(a->max_pairs >= 0 && a->max_pairs <= kMaxSlices * kPairsPerItem, "max_pairs out of range");
  OCC_REQUIRE(a->n_tables >= 0 && (a->n_tables == 0 || (a->table_off && a->table_H)), "table list missing");
  OCC_REQUIRE(a->pyr_tiles == 0 || a->pyr_off != nullptr, "pyr_off is NULL");
  OCC_REQUIRE(a->n_points >= 0, "n_points is negative");
  Workspace w;
  const int64_t need = ws_layout(a->T, a->F, total, a->SF, a->L, a->incl_len, a->pyr_tiles, a->bricks, a->max_pairs,
                                 (char *)a->workspace, &w);
  OCC_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= need, "workspace too small");
  const float vsf = (float)a->voxel_size;
  const float inv_vs = (a->flags & 32) ? 1.0f / vsf : 0.f;       // flag bit 5: torch-CUDA's x * (1 / vs)
  // shared-memory bitset of the crop kernel: as many words as the largest label slot needs (unknown: the maximum)
  const int smem_words = (a->max_label_slots > 0)
                             ? (int)std::min<int64_t>(a->max_label_slots / 32 + 2, kSmemBitWords) : kSmemBitWords;
  const bool f64_only = (a->flags & 1) != 0;
  const int chunk = f64_only ? kChunk : 32;
  const bool fast = !f64_only && a->F > 0;
  const bool cull = fast && a->pyr_tiles > 0 && !(a->flags & 2);
  const bool brick_cull = cull && !(a->flags & 16);
  // one memset: counters, crop flags, both bitsets, pair masks, table records
  OCC_CUDA(cudaMemsetAsync(w.zero_begin, 0, (size_t)(w.zero_end - w.zero_begin), stream));
  SideStream *side = nullptr;
  if (fast) {                                       // fork: sensor-side setup on the side stream
    if (side_stream(&side)) return 1;
    OCC_CUDA(cudaEventRecord(side->fork, stream));
    OCC_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    ProfScope ps(kProfSide, side->stream);
    const int64_t n_sens = a->SF * a->L;
    if (cull) {
      if (a->ri_tile_live) {                                 // dead tiles are skipped: both levels start from zero
        OCC_CUDA(cudaMemsetAsync(w.pyr, 0, 4 * (size_t)a->pyr_tiles, side->stream));
        OCC_CUDA(cudaMemsetAsync(w.pyr2, 0, 4 * 16 * (size_t)a->pyr_tiles, side->stream));
      }
      k_pyr_build<<<dim3((unsigned)n_sens, kPyrRowGroups), 256, 0, side->stream>>>(a->sensors, a->ri_pool, a->pyr_off,
                                                                                   w.pyr, w.pyr2, a->ri_tile_live);
      OCC_KERNEL_OK("k_pyr_build");
    }
    if (a->n_tables > 0) {
      k_table_setup<<<dim3((unsigned)a->n_tables, kTabSplit), 256, 4 * 8192, side->stream>>>(
          a->n_tables, a->table_off, a->table_H, a->incl_pool, w.tabcoef, w.ub_pool, w.lut_pool);
      OCC_KERNEL_OK("k_table_setup");
    }
    OCC_CUDA(cudaEventRecord(side->join, side->stream));
  }
  if (a->F > 0 && a->n_points > 0) {
    ProfScope ps(kProfCrop, stream);
    // about one CTA per resident slot (148 SMs x OCC_FMINB): a second, partly filled wave would be the kernel's tail
    const int crop_chunk = (int)std::min<int64_t>(
        kCropChunkMax, align_up(std::max<int64_t>(ceil_div(a->n_points, (int64_t)kNumSMs * OCC_FMINB), 1), kCropChunkMin));
    k_crop_voxelize<<<(unsigned)ceil_div(a->n_points, crop_chunk), kFrameThreads, 4 * smem_words, stream>>>(
        a->F, a->n_points, crop_chunk, a->poses, a->points, a->point_stride, a->frame_pt_off, a->trk_frame_off, a->label_off, a->frame_trk,
        w.frame_kept, w.trk_flags, w.grids, w.bits, vsf, inv_vs, nullptr, nullptr, smem_words);
    OCC_KERNEL_OK("k_crop_voxelize");
  }
  // the redo pass: frames of corrected tracklets only (device-side list, usually short)
  auto launch_redo = [&](cudaStream_t rs) -> int {
    k_crop_voxelize<<<(unsigned)std::min<int64_t>(a->F, kNumSMs * 2), kFrameThreads, 4 * smem_words, rs>>>(
        a->F, a->n_points, 0, a->poses, a->points, a->point_stride, a->frame_pt_off, a->trk_frame_off, a->label_off, a->frame_trk,
        w.frame_kept, w.trk_flags, w.grids, w.bits, vsf, inv_vs, w.redo_list, w.counter + 2, smem_words);
    OCC_KERNEL_OK("k_crop_voxelize(redo)");
    return 0;
  };
  if (f64_only) {
    {
      ProfScope ps(kProfSetup, stream);
      k_tracklet_setup<<<(unsigned)ceil_div(a->T, 8), 256, 0, stream>>>(
          a->T, a->trk_frame_off, a->poses, a->frame_pt_off, w.frame_kept, a->label_off, vsf, inv_vs, chunk, w.grids,
          w.bits, w.trk_flags, w.redo_list, w.counter + 2, a->dims, a->sizes, a->status, a->n_unknown, a->n_steps);
      OCC_KERNEL_OK("k_tracklet_setup");
      if (a->F > 0 && launch_redo(stream)) return 1;
    }
    {
      ProfScope ps(kProfScan, stream);
      k_scan_chunks<<<1, 1024, 0, stream>>>(a->T, w.grids, w.chunk_off);
      OCC_KERNEL_OK("k_scan_chunks");
    }
    const int64_t max_items = ceil_div(total, chunk) + a->T;
    const int grid = (int)std::min<int64_t>(max_items, (int64_t)kNumSMs * 8);
    ProfScope ps(kProfVisibility, stream);
    k_visibility_f64<<<grid, kChunk, 0, stream>>>(a->T, a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors,
                                                  a->incl_pool, a->ri_pool, a->voxel_size, a->label_off, w.grids,
                                                  w.chunk_off, w.counter, w.bits, w.trk_flags, a->labels,
                                                  a->labels_u8, a->status, a->n_unknown, a->n_steps);
    OCC_KERNEL_OK("k_visibility_f64");
    return 0;
  }
  if (!fast) {                                      // no tracklet-frames at all: only the statuses are due
    k_tracklet_setup<<<(unsigned)ceil_div(a->T, 8), 256, 0, stream>>>(
        a->T, a->trk_frame_off, a->poses, a->frame_pt_off, w.frame_kept, a->label_off, vsf, inv_vs, chunk, w.grids,
        w.bits, w.trk_flags, w.redo_list, w.counter + 2, a->dims, a->sizes, a->status, a->n_unknown, a->n_steps);
    OCC_KERNEL_OK("k_tracklet_setup");
    return 0;
  }
  OCC_CUDA(cudaStreamWaitEvent(stream, side->join, 0));          // tables and pyramid are ready
  {
    // the tracklet set-up runs as the prologue of each tracklet's CTA (one launch and one dependency level fewer)
    ProfScope ps(kProfPairBuild, stream);
    SetupArgs su;
    su.frame_pt_off = a->frame_pt_off; su.frame_kept = w.frame_kept; su.label_off = a->label_off;
    su.vsf = vsf; su.inv_vs = inv_vs; su.bits = w.bits; su.redo_list = w.redo_list; su.redo_count = w.counter + 2;
    su.dims_out = a->dims; su.sizes_out = a->sizes; su.n_unknown = a->n_unknown; su.n_steps = a->n_steps;
    k_pair_build<<<(unsigned)a->T, 256, 0, stream>>>(a->T, a->L, su, a->trk_frame_off, a->brick_off, w.grids, w.trk_flags,
                                                     a->poses, a->frame_sf, a->sensors, w.tabcoef, w.lut_pool,
                                                     a->voxel_size, a->pyr_off, w.pyr, cull ? 1 : 0, w.pairs_c, w.hot,
                                                     w.item_map, (long long)w.bricks, w.n_slices, w.counter, a->status);
    OCC_KERNEL_OK("k_pair_build");
  }
  {   // the redo pass runs on the side stream, next to k_brick_cull, and joins before the ray-cast
    OCC_CUDA(cudaEventRecord(side->fork, stream));
    OCC_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    ProfScope ps(kProfSetup, side->stream);
    if (launch_redo(side->stream)) return 1;
    if (w.bricks > 0) {
      k_brick_unknown<<<(unsigned)std::min<int64_t>(ceil_div(w.bricks, 8), (int64_t)kNumSMs * 8), 256, 0, side->stream>>>(
          w.item_map, w.counter, w.hot, w.bits, w.unk_brick);
      OCC_KERNEL_OK("k_brick_unknown");
    }
    OCC_CUDA(cudaEventRecord(side->join, side->stream));
  }
  // The ray-cast: brick cull + visibility.  With more than one slice of pairs the two kernels are split by slice
  // and overlapped: the cull of slices >= 1 (latency-bound: one short dependent chain per thread) runs on a second
  // side stream next to the visibility pass over slice 0 (issue-bound), which only needs slice 0's mask bits.
  const long long queue_cap_used = (a->flags & 8) ? std::min<long long>(64, (long long)w.queue_cap) : (long long)w.queue_cap;
  const bool split = OCC_OVERLAP && brick_cull && w.bricks > 0 && w.n_slices > 1;
  auto launch_cull = [&](int s0, int s1, cudaStream_t cs) -> int {
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div((w.bricks + 31) / 32 * 32 * kPairsPerItem, 256), (int64_t)kNumSMs * 16);
    k_brick_cull<<<dim3(gx, (unsigned)(s1 - s0)), 256, 0, cs>>>(s0, w.item_map, (long long)w.bricks, w.counter, w.hot,
                                                                 w.pairs_c, w.lut_pool, a->pyr_off, w.pyr2, w.mask_words,
                                                                 w.pair_mask);
    OCC_KERNEL_OK("k_brick_cull");
    return 0;
  };
  auto launch_vis = [&](int s0, int s1, int ticket) -> int {
    VisArgs va;
    va.L = a->L; va.mask_words = w.mask_words; va.s_lo = s0; va.s_hi = s1; va.ticket = ticket; va.pad = 0;
    va.bricks_total = (long long)w.bricks; va.queue_cap = queue_cap_used; va.vs = a->voxel_size;
    va.trk_frame_off = a->trk_frame_off; va.poses = a->poses; va.frame_sf = a->frame_sf; va.sensors = a->sensors;
    va.incl_pool = a->incl_pool; va.ri_pool = a->ri_pool; va.grids = w.grids; va.counter = w.counter;
    va.bits = w.bits; va.free_brick = w.free_brick; va.unk_brick = w.unk_brick; va.pair_mask = w.pair_mask; va.item_map = w.item_map;
    va.hot = w.hot; va.pairs = w.pairs_c; va.lut_pool = w.lut_pool; va.queue = w.queue; va.n_steps = a->n_steps;
    const int grid = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(w.bricks * (s1 - s0), 1), kFastWarps),
                                            (int64_t)kNumSMs * OCC_MINB);
    k_visibility<<<grid, 32 * kFastWarps, 0, stream>>>(va);
    OCC_KERNEL_OK("k_visibility");
    return 0;
  };
  if (split) {
    OCC_CUDA(cudaEventRecord(side->fork2, stream));           // after k_pair_build
    OCC_CUDA(cudaStreamWaitEvent(side->stream2, side->fork2, 0));
    if (launch_cull(1, w.n_slices, side->stream2)) return 1;
    OCC_CUDA(cudaEventRecord(side->join2, side->stream2));
  }
  if (brick_cull && w.bricks > 0) {
    ProfScope ps(kProfBrickCull, stream);
    if (launch_cull(0, split ? 1 : w.n_slices, stream)) return 1;
  }
  OCC_CUDA(cudaStreamWaitEvent(stream, side->join, 0));   // redo pass done: bits and flags final
  {
    ProfScope ps(kProfVisibility, stream);                  // (split: slice 0, the wait for the other culls, the rest)
    if (launch_vis(0, split ? 1 : w.n_slices, 0)) return 1;
    if (split) {
      OCC_CUDA(cudaStreamWaitEvent(stream, side->join2, 0));
      if (launch_vis(1, w.n_slices, 1)) return 1;
    }
  }
  {
    ProfScope ps(kProfRecheck, stream);
    k_visibility_recheck<<<kNumSMs * 4, 256, 0, stream>>>(a->L, a->trk_frame_off, a->poses, a->frame_sf, a->sensors,
                                                          a->incl_pool, a->ri_pool, a->voxel_size, w.grids, w.hot,
                                                          w.counter, w.queue, queue_cap_used, w.free_brick, a->n_steps);
    OCC_KERNEL_OK("k_visibility_recheck");
  }
  {
    ProfScope ps(kProfLabels, stream);
    k_labels<<<dim3((unsigned)a->T, 8), 256, 0, stream>>>(w.hot, w.grids, w.trk_flags, a->label_off, w.bits,
                                                          w.free_brick, a->labels, a->labels_u8, a->n_unknown,
                                                          a->status);
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
  ws_layout(a->T, a->F, total, a->SF, a->L, a->incl_len, a->pyr_tiles, a->bricks, a->max_pairs, (char *)a->workspace, &w);
  const float vsf = (float)a->voxel_size;
  k_frame_points<<<(unsigned)a->F, 256, 0, stream>>>(a->poses, a->points, a->point_stride, a->frame_pt_off,
                                                     a->frame_trk, w.grids, a->status, vsf,
                                                     (a->flags & 32) ? 1.0f / vsf : 0.f, loc_out, q_out);
  OCC_KERNEL_OK("k_frame_points");
  return 0;
}

extern "C" int occb200_annotate_queue_stats(const occb200_annotate_args_t *a, int64_t total, int64_t *out_host,
                                            void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Workspace w;
  ws_layout(a->T, a->F, total, a->SF, a->L, a->incl_len, a->pyr_tiles, a->bricks, a->max_pairs, (char *)a->workspace, &w);
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
  // max_err_host[0]: full-quadrant path (bound kAtanWide), max_err_host[1]: narrow path (bound kAtanNarrow)
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned long long *d = nullptr;
  OCC_CUDA(cudaMallocAsync((void **)&d, 16, stream));
  OCC_CUDA(cudaMemsetAsync(d, 0, 16, stream));
  k_selftest_atan2<<<kNumSMs * 4, 256, 0, stream>>>(n, seed, 0, d);
  OCC_KERNEL_OK("k_selftest_atan2");
  k_selftest_atan2<<<kNumSMs * 4, 256, 0, stream>>>(n, seed, 1, d + 1);
  OCC_KERNEL_OK("k_selftest_atan2");
  OCC_CUDA(cudaMemcpyAsync(max_err_host, d, 16, cudaMemcpyDeviceToHost, stream));
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
