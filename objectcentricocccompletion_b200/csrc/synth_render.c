/* synth_render.c -- the per-pixel loops of the synthetic scene generator (synth.py), in C with OpenMP.
 *
 * Test/benchmark infrastructure, not part of the product path: it renders the analytic object shells into the
 * range images of a synthetic segment and collects the candidate returns of each tracklet-frame.  Every
 * expression mirrors the NumPy reference implementation in synth.py (`_Renderer.add_object`,
 * `_Renderer.candidate_points`, `_slab_hit`, `_shell_hit`, `_to_box_frame`) operation by operation in f64 and is
 * compiled with -ffp-contract=off, so both generators produce the same bits (tests/test_synth_fast.py).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  int32_t H, W;
  double max_range;
  double o[3];          /* sensor origin, vehicle frame */
  const double *dirs;   /* [H, W, 3] unit directions, vehicle frame */
  double *img;          /* [B, H, W] f64 working images (render) */
  const float *img32;   /* [B, H, W] final f32 images (candidates) */
} synth_lidar_t;

/* per (object, LiDAR) pixel window: rows r_lo..r_hi, per frame columns (col0[b] + j) mod W, j < ncols */
typedef struct {
  int32_t ok;           /* 0: this LiDAR does not see the object */
  int32_t r_lo, r_hi, ncols;
} synth_win_t;

static inline double slab_hit(const double o[3], const double d[3], const double lo[3], const double hi[3]) {
  double tn = -INFINITY, tf = INFINITY;
  for (int k = 0; k < 3; ++k) {
    const double dk = fabs(d[k]) < 1e-12 ? 1e-12 : d[k];
    const double inv = 1.0 / dk;
    const double t0 = (lo[k] - o[k]) * inv, t1 = (hi[k] - o[k]) * inv;
    const double a = t0 < t1 ? t0 : t1, b = t0 > t1 ? t0 : t1;
    if (k == 0) { tn = a; tf = b; }
    else { tn = tn > a ? tn : a; tf = tf < b ? tf : b; }
  }
  const double tn0 = tn > 0.0 ? tn : 0.0;
  if (!(tf >= tn0)) return INFINITY;
  return tn > 0 ? tn : tf;
}

/* shape = {shrink, body_h, cab_w, cab_back, cab_front} */
static inline double shell_hit(const double o[3], const double d[3], const double size[3], const double *sh) {
  const double w = size[0], l = size[1], h = size[2];
  const double hb = h * sh[1];
  const double lo1[3] = {-0.5 * w * sh[0], -0.5 * l * sh[0], 0.0 + 0.02};
  const double hi1[3] = {0.5 * w * sh[0], 0.5 * l * sh[0], hb};
  const double lo2[3] = {-0.5 * w * sh[2], -0.5 * l * sh[3], hb};
  const double hi2[3] = {0.5 * w * sh[2], 0.5 * l * sh[4], h * sh[0]};
  const double a = slab_hit(o, d, lo1, hi1), b = slab_hit(o, d, lo2, hi2);
  return a < b ? a : b;
}

static inline int64_t wrap_col(int64_t c, int64_t W) {
  c %= W;
  return c < 0 ? c + W : c;
}

/* Render nobj shells into the images of nl LiDARs.
 *   boxes [nobj, B, 7] f64, cs [nobj, B, 2] = cos/sin(rz) as NumPy evaluated them, size_true [nobj, 3],
 *   shape [nobj, 5], frames [nobj, B] (0: the object is absent from that frame),
 *   win [nobj, nl], col0 [nobj, nl, B] */
void synth_render_objects(const synth_lidar_t *lidars, int nl, int B, int nobj, const double *boxes, const double *cs,
                          const double *size_true, const double *shape, const uint8_t *frames,
                          const synth_win_t *win, const int64_t *col0) {
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int li = 0; li < nl; ++li) {
    for (int b = 0; b < B; ++b) {
      const synth_lidar_t *ld = &lidars[li];
      const int H = ld->H, W = ld->W;
      double *img = ld->img + (int64_t)b * H * W;
      for (int ob = 0; ob < nobj; ++ob) {
        const synth_win_t wn = win[(int64_t)ob * nl + li];
        if (!wn.ok || !frames[(int64_t)ob * B + b]) continue;
        const double *box = boxes + ((int64_t)ob * B + b) * 7;
        const double c = cs[((int64_t)ob * B + b) * 2], s = cs[((int64_t)ob * B + b) * 2 + 1];
        const double tx = ld->o[0] - box[0], ty = ld->o[1] - box[1], tz = ld->o[2] - box[2];
        const double ol[3] = {tx * c - ty * s, tx * s + ty * c, tz};
        const int64_t c0 = col0[((int64_t)ob * nl + li) * B + b];
        for (int r = wn.r_lo; r <= wn.r_hi; ++r) {
          for (int j = 0; j < wn.ncols; ++j) {
            const int64_t col = wrap_col(c0 + j, W);
            const double *d = ld->dirs + ((int64_t)r * W + col) * 3;
            const double dl[3] = {d[0] * c - d[1] * s, d[0] * s + d[1] * c, d[2]};
            double t = shell_hit(ol, dl, size_true + 3 * (int64_t)ob, shape + 5 * (int64_t)ob);
            if (!(t <= ld->max_range)) continue;          /* inf: no change */
            double *px = img + (int64_t)r * W + col;
            const double cur = *px > 0 ? *px : INFINITY;
            const double nw = cur < t ? cur : t;
            *px = isfinite(nw) ? nw : 0.0;
          }
        }
      }
    }
  }
}

/* Candidate returns of every (object, frame): pixels of the rendered f32 images whose return lies inside the box
 * enlarged by pad.  Pass 1 (points == NULL) writes counts [nobj, B]; pass 2 writes the points (f32 xyz) of
 * (object, frame) at offsets[ob * B + b], LiDAR-major, window row-major -- the order of the NumPy version. */
void synth_candidate_points(const synth_lidar_t *lidars, int nl, int B, int nobj, const double *boxes,
                            const double *cs, double pad, const synth_win_t *win, const int64_t *col0,
                            int64_t *counts, const int64_t *offsets, float *points) {
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
  for (int ob = 0; ob < nobj; ++ob) {
    for (int b = 0; b < B; ++b) {
      const double *box = boxes + ((int64_t)ob * B + b) * 7;
      const double c = cs[((int64_t)ob * B + b) * 2], s = cs[((int64_t)ob * B + b) * 2 + 1];
      const double lim_x = 0.5 * box[3] + pad, lim_y = 0.5 * box[4] + pad;
      const double z_lo = -0.5 * pad, z_hi = box[5] + 0.5 * pad;
      int64_t n = 0;
      float *out = points ? points + 3 * offsets[(int64_t)ob * B + b] : 0;
      for (int li = 0; li < nl; ++li) {
        const synth_win_t wn = win[(int64_t)ob * nl + li];
        if (!wn.ok) continue;
        const synth_lidar_t *ld = &lidars[li];
        const int H = ld->H, W = ld->W;
        const float *img = ld->img32 + (int64_t)b * H * W;
        const int64_t c0 = col0[((int64_t)ob * nl + li) * B + b];
        for (int r = wn.r_lo; r <= wn.r_hi; ++r) {
          for (int j = 0; j < wn.ncols; ++j) {
            const int64_t col = wrap_col(c0 + j, W);
            const double rr = (double)img[(int64_t)r * W + col];
            if (!(rr > 0)) continue;
            const double *d = ld->dirs + ((int64_t)r * W + col) * 3;
            const double p0 = ld->o[0] + d[0] * rr, p1 = ld->o[1] + d[1] * rr, p2 = ld->o[2] + d[2] * rr;
            const double tx = p0 - box[0], ty = p1 - box[1], tz = p2 - box[2];
            const double l0 = tx * c - ty * s, l1 = tx * s + ty * c;
            if (fabs(l0) <= lim_x && fabs(l1) <= lim_y && tz >= z_lo && tz <= z_hi) {
              if (out) {
                out[3 * n] = (float)p0;
                out[3 * n + 1] = (float)p1;
                out[3 * n + 2] = (float)p2;
              }
              ++n;
            }
          }
        }
      }
      if (counts) counts[(int64_t)ob * B + b] = n;
    }
  }
}

/* [n] f64 -> f32 (the final cast of the rendered images), threaded */
void synth_f64_to_f32(const double *src, float *dst, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

/* dst[b] = base for b < B: the per-frame copies of the static scene, threaded */
void synth_repeat(const double *base, double *dst, int64_t n, int B) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) memcpy(dst + (int64_t)b * n, base, (size_t)n * sizeof(double));
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank: the generator takes its share of the host cores explicitly */
void synth_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
