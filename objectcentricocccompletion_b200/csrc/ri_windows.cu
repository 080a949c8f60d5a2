// ri_windows.cu -- upload only the pixels of the range images the visibility test can read.
//
// The reference loads the five whole range images of every frame (tools/occ/occ_annotate.py:502-533: 2.6 MB per
// frame, 104 MB for the 40 frames of one segment) and reads, per tracklet, a window of a few hundred pixels of
// each: the pixels the voxel centres of the tracklet's grid project to (:141-201, :541-547).  On the B200 path
// the host -> device copy of those images was the whole end-to-end step (134 MB over PCIe against 0.18 ms of
// kernels).  Two ways to move only the windows, both filling the same dense, zero-initialised device pool (the
// kernels of annotate.cu are unchanged: they index the dense pool):
//
//   pull  (occb200_pull_windows)  the range images stay in PINNED host memory in the pool's layout; the device
//         marks the 32-byte blocks the batch can read (k_window_mark: footprints of <= 0.8 m sub-boxes of each
//         tracklet's centre box) and reads exactly those over PCIe (k_window_pull: unified addressing).  No host
//         work per step, no staging copy.
//   host  (occb200_host_window_mark + occb200_host_gather_blocks + occb200_scatter_blocks)  for images in pageable
//         memory: the host marks the blocks with the same footprint code (one box per tracklet-frame), gathers
//         them from the source arrays into one staging buffer, the device scatters them to their place.
//
// The centre box.  In the box frame every voxel centre is idx*vs + min_bound + vs/2 with 0 <= idx < dims =
// ceil(size / vs) and min_bound = (-sx/2, -sy/2, 0) (occ_annotate.py:414-423, 467-471), size <= S = the max of the box
// size over ALL frames of the tracklet (the true size is the max over the frames that keep a point, :111-112,
// :132-133).  So the centres lie in  [-S/2, S/2 + vs] x [-S/2, S/2 + vs] x [0, Sz + vs]  (+ 1 cm).  The chain box frame
// -> ego -> sensor (:490-499, :161-164) is affine, so a (sub-)box maps to a parallelepiped whose corner extremes
// bracket range, height and azimuth (sub_footprint below).
//
// Why the labels cannot change.  Every pixel the exact test of a real voxel centre can read lies inside a marked
// block.  What lies outside keeps its previous content (zeros, or an earlier upload of the same pool): only the
// max-pyramid and the discarded tests of padding lanes ever look at it.  A pyramid tile that holds such pixels can
// only make a cull LESS likely than the truth restricted to the window requires -- or more likely, but then every
// in-window pixel is below the cull's bound and the culled tests would all have failed.  (n_steps, the count of
// evaluated tests, may differ from a whole-image upload for that reason; labels, dims and statuses cannot.)
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace occb200 {

constexpr int kBlk = 16;      // floats per block: 64 bytes = two sectors

__global__ void __launch_bounds__(256)
k_scatter_blocks(const float4 *__restrict__ blocks, const uint32_t *__restrict__ block_idx, long long n_blocks,
                 float *__restrict__ ri_pool, long long ri_len) {
  // four threads per block, one 16-byte piece each: coalesced reads, 64-byte aligned writes
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks * 4;
       i += (long long)gridDim.x * blockDim.x) {
    const long long dst = (long long)__ldg(block_idx + (i >> 2)) * kBlk + 4 * (i & 3);
    const float4 v = ld_stream4(blocks + i);
    if (dst + 4 <= ri_len) {
      *reinterpret_cast<float4 *>(ri_pool + dst) = v;
    } else {                                         // last, partial block of the pool
      const float e[4] = {v.x, v.y, v.z, v.w};
      for (int k = 0; k < 4; ++k)
        if (dst + k < ri_len) ri_pool[dst + k] = e[k];
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Device-side windows: the GPU derives the footprint itself and PULLS the blocks it needs straight from the
// pinned host copy of the range images (unified addressing: a kernel may read pinned host memory over PCIe).
// No host geometry, no host gather, no staging buffer: the host only keeps the images where the loader put them.
//
// k_window_mark   one thread per (tracklet-frame, LiDAR, sub-box): the tracklet's centre box (file header) is cut
//                 into sub-boxes of <= kSubEdge metres; each sub-box maps to a
//                 convex body in the sensor frame whose pixel footprint is bracketed from its 8 corners exactly as
//                 k_brick_cull does for a brick (annotate.cu): range r_lo <= |p| <= r_hi, z extremes at corners,
//                 sin(inc) = z / |p|, azimuth half-width from cross / dot sums.  The union of the sub-box footprints
//                 hugs the object's outline (one ball around the whole box marks 2x the pixels).  f32 with
//                 generous padding: 1e-4 rad, 1 mm + 1e-5 d, one row, two columns; any doubt marks whole rows /
//                 the whole image.  Sets one bit per 32-byte block (8 floats) of the pool.
// k_window_pull   one warp per 32 mask bits (1 KB of the pool): the marked blocks are read from the host pool with
//                 16-byte loads (adjacent marked blocks coalesce into full PCIe requests) and stored at the same
//                 offset of the device pool.
// ---------------------------------------------------------------------------------------------
constexpr int kPullBlk = 8;            // floats per block of the device-side path: one 32-byte sector
constexpr float kSubEdge = 0.8f;       // metres

struct WinArgs {
  int T, L;
  long long SF, ri_len;
  const int64_t *trk_frame_off;
  const occb200_pose_t *poses;
  const int32_t *frame_sf;
  const occb200_sensor_t *sensors;
  const float *incl_pool;
  const float *trk_smax;   // [T,3] max box size over ALL frames of the tracklet (the grid's upper bound)
  float vs;
  uint32_t *mask;
};

__device__ __forceinline__ void mark_blocks(uint32_t *__restrict__ mask, long long a, long long b) {   // floats [a, b]
  const long long b0 = a / kPullBlk, b1 = b / kPullBlk;
  for (long long w = b0 >> 5; w <= (b1 >> 5); ++w) {
    const int lo = (w == (b0 >> 5)) ? (int)(b0 & 31) : 0, hi = (w == (b1 >> 5)) ? (int)(b1 & 31) : 31;
    const uint32_t bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
    if ((__ldg(mask + w) & bits) != bits) atomicOr(mask + w, bits);     // neighbours have usually set them already
  }
}

// number of table entries above x (descending table) -> nearest row is that or the one before: the caller pads by one
__host__ __device__ inline int rows_above(const float *__restrict__ tab, int H, float x) {
  int lo = 0, hi = H;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (tab[mid] > x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The tracklet's centre box cut into n[0] x n[1] x n[2] sub-boxes (box frame): lo = its low corner, ext = its extent.
struct SubGrid {
  float lo[3], ext[3];
  int n[3];
  bool sane;
  int nsub;
};

__host__ __device__ inline SubGrid sub_grid(const float *__restrict__ smax3, float vs, float edge = kSubEdge) {
  SubGrid g;
  const float m = 0.01f;
  g.lo[0] = -0.5f * smax3[0] - m; g.lo[1] = -0.5f * smax3[1] - m; g.lo[2] = -m;
  g.ext[0] = smax3[0] + vs + 2.f * m; g.ext[1] = smax3[1] + vs + 2.f * m; g.ext[2] = smax3[2] + vs + 2.f * m;
  g.sane = true;
  for (int k = 0; k < 3; ++k)
    if (!(g.ext[k] > 0.f && g.ext[k] < 1e4f)) g.sane = false;
  for (int k = 0; k < 3; ++k) {
    const int c = (int)ceilf(g.ext[k] / edge);
    g.n[k] = g.sane ? (c < 1 ? 1 : (c > 64 ? 64 : c)) : 1;     // absurd sizes: one item per pair marks the whole image
  }
  g.nsub = g.n[0] * g.n[1] * g.n[2];
  return g;
}

// Pixel footprint of sub-box `sub` through one (frame, LiDAR): rows [r0, r1], columns [c_lo, c_hi] (possibly beyond
// [0, W): taken modulo W by the caller; all_cols: every column).  Compiled for the device (k_window_mark) and for the
// host (occb200_host_window_mark, the CPU test hook of the same code).
struct Footprint {
  int r0, r1;
  long long c_lo, c_hi;
  bool all_cols;
};

__host__ __device__ inline Footprint sub_footprint(const SubGrid &g, int sub, const occb200_pose_t &ps,
                                                   const occb200_sensor_t &sn, const float *__restrict__ tab) {
  const int H = sn.H, W = sn.W;
  Footprint fp;
  fp.r0 = 0; fp.r1 = H - 1; fp.c_lo = 0; fp.c_hi = W - 1; fp.all_cols = true;
  if (!g.sane) return fp;
  bool all_rows = sn.incl_mono != -1;
  // sub-box centre / half extents in the box frame
  const int s0 = sub / (g.n[1] * g.n[2]), rem = sub - s0 * (g.n[1] * g.n[2]);
  const int s1 = rem / g.n[2], s2 = rem - s1 * g.n[2];
  const float h0 = 0.5f * g.ext[0] / g.n[0], h1 = 0.5f * g.ext[1] / g.n[1], h2 = 0.5f * g.ext[2] / g.n[2];
  const float x0 = g.lo[0] + (2 * s0 + 1) * h0, x1 = g.lo[1] + (2 * s1 + 1) * h1, x2 = g.lo[2] + (2 * s2 + 1) * h2;
  // box frame -> ego (occ_annotate.py:490-499): ego = (x0 c + x1 s, -x0 s + x1 c, x2) + origin; ego -> sensor (:161-164)
  const float c = ps.cos_p, s = ps.sin_p;
  const float *v = sn.v2l;
  float M[9];                                   // sensor <- box frame: M = V Rm, Rm = [[c, s, 0], [-s, c, 0], [0, 0, 1]]
  for (int r = 0; r < 3; ++r) {
    M[3 * r] = v[4 * r] * c - v[4 * r + 1] * s;
    M[3 * r + 1] = v[4 * r] * s + v[4 * r + 1] * c;
    M[3 * r + 2] = v[4 * r + 2];
  }
  float pc[3];
  for (int r = 0; r < 3; ++r)
    pc[r] = M[3 * r] * x0 + M[3 * r + 1] * x1 + M[3 * r + 2] * x2 + v[4 * r] * ps.box[0] + v[4 * r + 1] * ps.box[1] +
            v[4 * r + 2] * ps.box[2] + v[4 * r + 3];
  // the 8 corners are pc +- a +- b +- c: every corner extreme below is a sum of absolute values
  const float ax = h0 * M[0], ay = h0 * M[3], az = h0 * M[6];
  const float bx = h1 * M[1], by = h1 * M[4], bz = h1 * M[7];
  const float cx = h2 * M[2], cy = h2 * M[5], cz = h2 * M[8];
  const float R = sqrtf(h0 * h0 + h1 * h1 + h2 * h2) * 1.001f + 1e-3f;
  const float rho2 = pc[0] * pc[0] + pc[1] * pc[1], d2 = rho2 + pc[2] * pc[2];
  const float rho = sqrtf(rho2), d = sqrtf(d2);
  if (!(d > 1.25f * R) || !(d < 1e6f)) return fp;           // the sensor sits in / next to the sub-box: whole image
  const float ux = pc[0] / d, uy = pc[1] / d, uz = pc[2] / d;
  const float slack = 1e-3f + 1e-5f * d;
  // |p| over the body: >= min over corners of p.u (linear), <= |pc| + half diagonal
  const float r_lo = d - (fabsf(ax * ux + ay * uy + az * uz) + fabsf(bx * ux + by * uy + bz * uz) +
                          fabsf(cx * ux + cy * uy + cz * uz)) - slack;
  const float r_hi = d + R + slack;
  const float zext = fabsf(az) + fabsf(bz) + fabsf(cz) + slack;
  const float zmin = pc[2] - zext, zmax = pc[2] + zext;
  if (!(r_lo > 0.f)) return fp;
  if (!all_rows) {
    // sin(inc) = z / |p| (:165-166), bracketed with the sign of z
    const float s_hi = fminf(fmaxf(zmax / (zmax > 0.f ? r_lo : r_hi), -1.f), 1.f);
    const float s_lo = fminf(fmaxf(zmin / (zmin > 0.f ? r_hi : r_lo), -1.f), 1.f);
    const float inc_hi = asinf(s_hi) + 1e-4f, inc_lo = asinf(s_lo) - 1e-4f;
    // nearest entry to x (:168-173) is entry k - 1 or k, k = rows_above(x); one more row of padding on each side
    const int k_hi = rows_above(tab, H, inc_hi), k_lo = rows_above(tab, H, inc_lo);
    fp.r0 = k_hi - 2 > 0 ? k_hi - 2 : 0;
    fp.r1 = k_lo + 1 < H - 1 ? k_lo + 1 : H - 1;
  }
  // azimuth relative to the centre's: tan = cross(pc, v) / dot(pc, v) in the xy plane; over the corners
  // |cross| <= sum |cross(pc, e)| and dot >= rho^2 - sum |dot(pc, e)|
  const float crs = fabsf(pc[0] * ay - pc[1] * ax) + fabsf(pc[0] * by - pc[1] * bx) + fabsf(pc[0] * cy - pc[1] * cx);
  const float dmin = rho2 - (fabsf(pc[0] * ax + pc[1] * ay) + fabsf(pc[0] * bx + pc[1] * by) + fabsf(pc[0] * cx + pc[1] * cy));
  if (!(rho > 1.05f * R) || !(dmin > 0.125f * rho2)) return fp;     // over / under the sensor: every column
  const float dphi = atanf(crs / dmin) + 1e-4f;
  const float kc = (float)W * 0.15915494309189535f;
  const float azm = atan2f(pc[1], pc[0]) + sn.azc;
  // colf = (W - 0.5) - (az + pi) / (2 pi) W  (:187-189); f32 evaluation error << the two columns of padding
  const float cf = ((float)W - 0.5f) - (azm + 3.14159265358979f) * kc;
  const float cf_lo = cf - dphi * kc, cf_hi = cf + dphi * kc;
  if (cf_hi - cf_lo + 8.f < (float)W) {
    fp.c_lo = (long long)floorf(cf_lo) - 2;
    fp.c_hi = (long long)ceilf(cf_hi) + 2;
    fp.all_cols = false;
  }
  return fp;
}

// Marks the footprint in a block mask through `mark(first float, last float)` of the pool.
template <typename Mark>
__host__ __device__ inline void mark_footprint(const Footprint &fp, long long img, int W, Mark mark) {
  if (fp.all_cols || fp.c_hi - fp.c_lo + 1 >= W) {
    mark(img + (long long)fp.r0 * W, img + (long long)fp.r1 * W + W - 1);
    return;
  }
  const long long a0 = ((fp.c_lo % W) + W) % W, len = fp.c_hi - fp.c_lo + 1;
  for (int r = fp.r0; r <= fp.r1; ++r) {
    const long long row = img + (long long)r * W;
    mark(row + a0, row + (a0 + len < W ? a0 + len : W) - 1);
    if (a0 + len > W) mark(row, row + (a0 + len - W) - 1);
  }
}

// The pyramid tiles (8 rows x 32 columns, annotate.cu's kTileR x kTileC) that hold a marked block, derived from the
// block mask once it is complete: one thread per tile looks at the mask bits of its 8 row segments (<= 5 blocks each).
// A tile that stays 0 holds no pixel a visibility test can read.  (Marking the tiles from inside k_window_mark made
// hundreds of threads store to the same bytes: the end-to-end step lost 4 %.)
__global__ void __launch_bounds__(256)
k_tiles_from_mask(const occb200_sensor_t *__restrict__ sensors, const int64_t *__restrict__ pyr_off,
                  const uint32_t *__restrict__ mask, long long ri_len, uint8_t *__restrict__ tile_live) {
  const int e = blockIdx.x;
  const occb200_sensor_t &sn = sensors[e];
  const int H = sn.H, W = sn.W;
  if (H < 1 || W < 1) return;
  const int ntr = (H + 7) / 8, ntc = (W + 31) / 32;
  uint8_t *live = tile_live + pyr_off[e];
  const bool inside = sn.ri_off >= 0 && sn.ri_off + (long long)H * W <= ri_len;
  for (int tile = blockIdx.y * blockDim.x + threadIdx.x; tile < ntr * ntc; tile += gridDim.y * blockDim.x) {
    const int tr = tile / ntc, tc = tile - tr * ntc;
    bool any = !inside;                                            // (an image outside the pool is never marked: build it)
    for (int r = 0; r < 8 && !any; ++r) {
      const int row = tr * 8 + r;
      if (row >= H) break;
      const long long o0 = sn.ri_off + (long long)row * W + tc * 32;
      const long long o1 = o0 + min(31, W - 1 - tc * 32);
      const long long b0 = o0 / kPullBlk, b1 = o1 / kPullBlk;       // <= 5 blocks: one or two mask words
      for (long long w = b0 >> 5; w <= (b1 >> 5); ++w) {
        const int lo = (w == (b0 >> 5)) ? (int)(b0 & 31) : 0, hi = (w == (b1 >> 5)) ? (int)(b1 & 31) : 31;
        const uint32_t bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        any = any || (__ldg(mask + w) & bits) != 0u;
      }
    }
    live[tile] = any ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) k_window_mark(const WinArgs a) {
  const int t = blockIdx.y;
  const int64_t f0 = a.trk_frame_off[t], f1 = a.trk_frame_off[t + 1];
  const int B = (int)(f1 - f0);
  if (B <= 0) return;
  const SubGrid g = sub_grid(a.trk_smax + 3 * t, a.vs);
  const long long items = (long long)B * a.L * g.nsub;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int sub = (int)(it % g.nsub);
    const int pr = (int)(it / g.nsub);
    const int i = pr / a.L, l = pr - i * a.L;
    const long long sf = a.frame_sf[f0 + i];
    if (sf < 0 || sf >= a.SF) continue;
    const occb200_sensor_t &sn = a.sensors[sf * a.L + l];
    const long long img = sn.ri_off;
    if (sn.H < 1 || sn.W < 1 || img < 0 || img + (long long)sn.H * sn.W > a.ri_len) continue;
    const Footprint fp = sub_footprint(g, sub, a.poses[f0 + i], sn, a.incl_pool + sn.incl_off);
    uint32_t *mask = a.mask;
    mark_footprint(fp, img, sn.W, [mask](long long x, long long y) { mark_blocks(mask, x, y); });
  }
}

__global__ void __launch_bounds__(256)
k_window_pull(const uint32_t *__restrict__ mask, long long n_words, const float *__restrict__ src,
              float *__restrict__ dst, long long ri_len, unsigned long long *__restrict__ pulled) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long count = 0;
  for (long long w = warp; w < n_words; w += nwarps) {
    const uint32_t bits = __ldg(mask + w);
    if (!bits) continue;
    count += __popc(bits);
    // the word's 32 blocks = 256 floats = 64 float4: lane takes float4 number lane and lane + 32
    const long long base = w * 32 * kPullBlk;
    float4 v[2];
    bool on[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int q = lane + 32 * j;                    // float4 index inside the word; block = q >> 1
      const long long e = base + 4 * q;
      on[j] = ((bits >> (q >> 1)) & 1u) && e + 4 <= ri_len;
      if (on[j]) v[j] = ld_stream4(reinterpret_cast<const float4 *>(src + e));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (on[j]) *reinterpret_cast<float4 *>(dst + base + 4 * (lane + 32 * j)) = v[j];
  }
  if (pulled && lane == 0 && count) atomicAdd(pulled, count);
}

}  // namespace occb200

using namespace occb200;

// HOST.  Compacts an 8-float block mask into the ascending list of 16-float blocks (64 bytes, the unit of the gather /
// scatter pair) that hold a marked block; returns their number.  out has room for ceil(n8 / 2) entries.
extern "C" int64_t occb200_host_mask_to_blocks(const uint8_t *mask8, int64_t n8, uint32_t *out) {
  // two passes over chunks of 64 K mask bytes (OpenMP): count the pairs that hold a marked block, prefix, write.
  // (The one-pass scalar loop was 2.6 ms of a 6.5 ms one-shot call for the 3.2 M mask bytes of one segment.)
  const int64_t npair = (n8 + 1) / 2;                 // pair k = mask bytes 2k, 2k+1 (the last one may be half)
  const int64_t chunk = 32768;                        // pairs per chunk
  const int64_t nchunk = ceil_div(npair, chunk);
  if (nchunk == 0) return 0;
  std::vector<int64_t> cnt(nchunk + 1, 0);
  auto marked = [&](int64_t k) -> bool { return mask8[2 * k] | ((2 * k + 1 < n8) ? mask8[2 * k + 1] : 0); };
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nchunk; ++c) {
    const int64_t k0 = c * chunk, k1 = std::min(npair, k0 + chunk);
    int64_t n = 0, k = k0;
    for (; k + 4 <= k1 && 2 * (k + 4) <= n8; k += 4) {             // 8 mask bytes at a time: most words are all zero
      uint64_t w;
      memcpy(&w, mask8 + 2 * k, 8);
      if (!w) continue;
      for (int j = 0; j < 4; ++j) n += ((w >> (16 * j)) & 0xffffu) ? 1 : 0;
    }
    for (; k < k1; ++k) n += marked(k) ? 1 : 0;
    cnt[c + 1] = n;
  }
  for (int64_t c = 0; c < nchunk; ++c) cnt[c + 1] += cnt[c];
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nchunk; ++c) {
    const int64_t k0 = c * chunk, k1 = std::min(npair, k0 + chunk);
    uint32_t *o = out + cnt[c];
    int64_t k = k0;
    for (; k + 4 <= k1 && 2 * (k + 4) <= n8; k += 4) {
      uint64_t w;
      memcpy(&w, mask8 + 2 * k, 8);
      if (!w) continue;
      for (int j = 0; j < 4; ++j)
        if ((w >> (16 * j)) & 0xffffu) *o++ = (uint32_t)(k + j);
    }
    for (; k < k1; ++k)
      if (marked(k)) *o++ = (uint32_t)k;
  }
  return cnt[nchunk];
}

// HOST.  Copies the listed blocks from the source arrays to staging[16 * i ..]: block k of the pool lies in part
// j = the last one with part_off[j] <= 16 k, at part_ptr[j] + (16 k - part_off[j]); floats beyond the part's end
// (a last, partial block) are zero.  block_idx ascending, part_off ascending multiples of 16.
extern "C" int occb200_host_gather_blocks(const uint32_t *block_idx, int64_t n_blocks, const int64_t *part_off,
                                          const int64_t *part_len, const float *const *part_ptr, int32_t n_parts,
                                          float *staging) {
  OCC_REQUIRE(n_blocks >= 0 && n_parts >= 0, "bad arguments");
  if (n_blocks == 0) return 0;
  OCC_REQUIRE(n_parts > 0, "blocks without source arrays");
  int bad = 0;
#pragma omp parallel
  {
    int j = 0;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n_blocks; ++i) {
      const int64_t a = (int64_t)block_idx[i] * kBlk;
      if (a < part_off[j]) j = 0;                    // (a thread's range starts anywhere)
      while (j + 1 < n_parts && part_off[j + 1] <= a) ++j;
      const int64_t rel = a - part_off[j];
      float *dst = staging + i * kBlk;
      if (rel < 0 || rel >= part_len[j]) { bad = 1; memset(dst, 0, 4 * kBlk); continue; }
      const int64_t n = std::min<int64_t>(kBlk, part_len[j] - rel);
      memcpy(dst, part_ptr[j] + rel, 4 * (size_t)n);
      if (n < kBlk) memset(dst + n, 0, 4 * (size_t)(kBlk - n));
    }
  }
  OCC_REQUIRE(!bad, "a block lies outside every source array");
  return 0;
}

// HOST.  dst[dst_off[i] .. dst_off[i] + bytes[i]) = src[i][0 .. bytes[i]) for n parts, in parallel (OpenMP, pieces of
// 256 KB): how the one-shot API moves its pageable inputs (candidate points of every tracklet, small fields) into
// ONE pinned staging buffer -- a pageable cudaMemcpy of the same 55 MB took 7.4 ms of a 12 ms call.
extern "C" int occb200_host_copy_parts(const void *const *src, const int64_t *bytes, const int64_t *dst_off, int32_t n,
                                       void *dst) {
  OCC_REQUIRE(n >= 0 && (n == 0 || (src && bytes && dst_off && dst)), "bad arguments");
  const int64_t piece = 256 * 1024;
  int64_t total = 0;
  for (int32_t i = 0; i < n; ++i) {
    OCC_REQUIRE(bytes[i] >= 0 && dst_off[i] >= 0, "negative size / offset");
    total += ceil_div(bytes[i], piece);
  }
  // piece -> (part, offset) by a prefix walk per thread chunk
  std::vector<int64_t> first(n + 1, 0);
  for (int32_t i = 0; i < n; ++i) first[i + 1] = first[i] + ceil_div(bytes[i], piece);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < total; ++q) {
    const int32_t i = (int32_t)(std::upper_bound(first.begin(), first.end(), q) - first.begin()) - 1;
    const int64_t o = (q - first[i]) * piece;
    const int64_t len = std::min<int64_t>(piece, bytes[i] - o);
    memcpy((char *)dst + dst_off[i] + o, (const char *)src[i] + o, (size_t)len);
  }
  return 0;
}

// DEVICE.  ri_pool[16 * block_idx[i] + j] = blocks[16 * i + j]: the second half of the windowed upload (blocks and
// block_idx are device copies of the staging buffer and the block list).  Asynchronous on `stream`.
extern "C" int occb200_scatter_blocks(const float *blocks, const uint32_t *block_idx, int64_t n_blocks, float *ri_pool,
                                      int64_t ri_len, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(n_blocks >= 0 && ri_len >= 0, "bad sizes");
  if (n_blocks == 0) return 0;
  OCC_REQUIRE(blocks && block_idx && ri_pool, "NULL argument");
  OCC_REQUIRE(((uintptr_t)blocks & 15) == 0 && ((uintptr_t)ri_pool & 15) == 0, "blocks and ri_pool must be 16-byte aligned");
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n_blocks * 4, 256), (int64_t)kNumSMs * 16);
  k_scatter_blocks<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(blocks), block_idx, (long long)n_blocks,
                                             ri_pool, (long long)ri_len);
  OCC_KERNEL_OK("k_scatter_blocks");
  return 0;
}

extern "C" int64_t occb200_window_mask_words(int64_t ri_len) {
  return ceil_div(ceil_div(ri_len, (int64_t)kPullBlk), 32) + 1;
}

// DEVICE.  The windowed upload without host work: marks the 32-byte blocks of the pool the batch can read
// (k_window_mark, from the device copies of the metadata) and pulls them from `ri_host` -- the PINNED host array
// that holds the range images in the layout of ri_pool (device-accessible under unified addressing) -- into
// ri_pool.  mask: device scratch of occb200_window_mask_words(ri_len) uint32.  *pulled_blocks (device, optional,
// caller-zeroed) accumulates the number of 32-byte blocks read over PCIe.  Asynchronous on `stream`.
extern "C" int occb200_pull_windows(const occb200_annotate_args_t *a, const float *trk_smax, const float *ri_host,
                                    float *ri_pool, int64_t ri_len, uint32_t *mask,
                                    unsigned long long *pulled_blocks, uint8_t *tile_live, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  OCC_REQUIRE(a != nullptr && ri_len >= 0, "bad arguments");
  if (a->T == 0 || a->F == 0 || ri_len == 0) return 0;
  OCC_REQUIRE(trk_smax && ri_host && ri_pool && mask, "NULL argument");
  OCC_REQUIRE(((uintptr_t)ri_host & 15) == 0 && ((uintptr_t)ri_pool & 15) == 0, "pools must be 16-byte aligned");
  OCC_REQUIRE(ri_len % kPullBlk == 0, "ri_len must be a multiple of 8 floats (pad the pool)");
  const int64_t n_words = occb200_window_mask_words(ri_len);
  OCC_CUDA(cudaMemsetAsync(mask, 0, 4 * (size_t)n_words, stream));
  OCC_REQUIRE(tile_live == nullptr || (a->pyr_off != nullptr && a->pyr_tiles > 0), "tile_live needs pyr_off / pyr_tiles");
  WinArgs w;
  w.T = a->T; w.L = a->L; w.SF = a->SF; w.ri_len = ri_len;
  w.trk_frame_off = a->trk_frame_off; w.poses = a->poses; w.frame_sf = a->frame_sf; w.sensors = a->sensors;
  w.incl_pool = a->incl_pool; w.trk_smax = trk_smax; w.vs = (float)a->voxel_size; w.mask = mask;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(32, ceil_div((int64_t)kNumSMs * 16, a->T)));
  k_window_mark<<<dim3(gx, (unsigned)a->T), 256, 0, stream>>>(w);
  OCC_KERNEL_OK("k_window_mark");
  const unsigned gp = (unsigned)std::min<int64_t>(ceil_div(n_words, 8), (int64_t)kNumSMs * 4);
  k_window_pull<<<gp, 256, 0, stream>>>(mask, (long long)n_words - 1, ri_host, ri_pool, (long long)ri_len, pulled_blocks);
  OCC_KERNEL_OK("k_window_pull");
  if (tile_live) {                                   // (behind the pull: nothing on this stream waits for it but the pyramid)
    k_tiles_from_mask<<<dim3((unsigned)(a->SF * a->L), 2), 256, 0, stream>>>(a->sensors, a->pyr_off, mask, (long long)ri_len,
                                                                             tile_live);
    OCC_KERNEL_OK("k_tiles_from_mask");
  }
  return 0;
}

// HOST: the footprint code of k_window_mark compiled for the CPU.  mask8[k] = 1 for every 8-float block some
// tracklet of the batch can read.  sub_edge <= 0: the device path's sub-boxes (the CPU tests compare the two, equal
// up to the last-ulp differences of the host's libm, far inside the padding); a large value: one box per
// tracklet-frame -- what the "host" upload mode uses (12 % more bytes than 0.8 m sub-boxes for 1/40 of the work).
extern "C" int occb200_host_window_mark(int32_t T, int32_t L, const int64_t *trk_frame_off, const occb200_pose_t *poses,
                                        const int32_t *frame_sf, const occb200_sensor_t *sensors, int64_t SF,
                                        const float *incl_pool, const float *trk_smax, double voxel_size,
                                        int64_t ri_len, uint8_t *mask8, float sub_edge) {
  OCC_REQUIRE(T >= 0 && L >= 1 && ri_len >= 0, "bad arguments");
  const float edge = sub_edge > 0.f ? sub_edge : kSubEdge;
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < T; ++t) {
    const int64_t f0 = trk_frame_off[t], f1 = trk_frame_off[t + 1];
    const SubGrid g = sub_grid(trk_smax + 3 * t, (float)voxel_size, edge);
    for (int64_t f = f0; f < f1; ++f) {
      const int64_t sf = frame_sf[f];
      if (sf < 0 || sf >= SF) continue;
      for (int l = 0; l < L; ++l) {
        const occb200_sensor_t &sn = sensors[sf * L + l];
        const long long img = sn.ri_off;
        if (sn.H < 1 || sn.W < 1 || img < 0 || img + (long long)sn.H * sn.W > ri_len) continue;
        for (int sub = 0; sub < g.nsub; ++sub) {
          const Footprint fp = sub_footprint(g, sub, poses[f], sn, incl_pool + sn.incl_off);
          mark_footprint(fp, img, sn.W, [mask8](long long x, long long y) {
            for (long long k = x / kPullBlk; k <= y / kPullBlk; ++k) mask8[k] = 1;
          });
        }
      }
    }
  }
  return 0;
}
