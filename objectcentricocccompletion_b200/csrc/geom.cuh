// geom.cuh -- device arithmetic that must reproduce the reference bit for bit.
//
// Every operation whose rounding matters is written with the round-to-nearest intrinsics
// (__fmul_rn, __fmaf_rn, __dmul_rn, __fma_rn, ...), which nvcc never contracts or reorders; the
// library is compiled with -fmad=false on top of that.  The operation ORDER follows what the
// reference's torch-CPU ops do (pinned by oracle/validate_oracle.py).
#pragma once
#include "common.cuh"

namespace occb200 {

// ---------------------------------------------------------------------------------------------
// A1: check_pt_in_box3d  (mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:24-49,
//                         same arithmetic as points_in_boxes_cpu.cpp:16-41)
// ---------------------------------------------------------------------------------------------
struct BoxTest {
  float cx, cy, czc;   // czc = float(double(z_bottom) + double(h)/2.0)
  float hx, hy, hz;    // half sizes: hx = l/2 bounds local_x, hy = w/2 bounds local_y, hz = h/2
  float cosa, sina;    // cos/sin(float(double(rz) + pi/2))
  int contract;        // 0: the CPU kernel's unfused products (points_in_boxes_cpu.cpp:27-28); 1: the FMA contraction
                       // nvcc applies to the reference's CUDA kernel (points_in_boxes_cuda.cu:31-32, SASS of the
                       // unmodified source built for sm_100a: local_x = fma(sx, cosa, -(sy * sina)),
                       // local_y = fma(sy, cosa, sx * sina))
};

__device__ __forceinline__ BoxTest make_box_test(const float *box7, float cosa, float sina, int contract = 0) {
  BoxTest b;
  b.contract = contract;
  b.cx = box7[0];
  b.cy = box7[1];
  // `cz += h / 2.0` : double add, rounded back to float (points_in_boxes_cuda.cu:41)
  b.czc = (float)__dadd_rn((double)box7[2], __ddiv_rn((double)box7[5], 2.0));
  // h/2.0, l/2.0, w/2.0 are exact in float, so the reference's double comparisons equal these
  b.hx = __fmul_rn(box7[4], 0.5f);
  b.hy = __fmul_rn(box7[3], 0.5f);
  b.hz = __fmul_rn(box7[5], 0.5f);
  b.cosa = cosa;
  b.sina = sina;
  return b;
}

// rot_angle = rz + M_PI/2 in double, rounded to float (points_in_boxes_cuda.cu:28)
__device__ __forceinline__ float box_rot_angle(float rz) {
  return (float)__dadd_rn((double)rz, 1.57079632679489661923);
}

__device__ __forceinline__ bool pt_in_box(const BoxTest &b, float x, float y, float z) {
  if (fabsf(__fsub_rn(z, b.czc)) > b.hz) return false;
  float sx = __fsub_rn(x, b.cx), sy = __fsub_rn(y, b.cy);
  float lx, ly;
  if (b.contract) {
    lx = __fmaf_rn(sx, b.cosa, -__fmul_rn(sy, b.sina));
    ly = __fmaf_rn(sy, b.cosa, __fmul_rn(sx, b.sina));
  } else {
    lx = __fadd_rn(__fmul_rn(sx, b.cosa), __fmul_rn(sy, -b.sina));
    ly = __fadd_rn(__fmul_rn(sx, b.sina), __fmul_rn(sy, b.cosa));
  }
  return (lx > -b.hx) & (lx < b.hx) & (ly > -b.hy) & (ly < b.hy);
}

// ---------------------------------------------------------------------------------------------
// A5: point_cloud_to_range_image_idx for one point, all-f64 (tools/occ/occ_annotate.py:141-201)
// ---------------------------------------------------------------------------------------------
struct SensorView {   // what one (frame, LiDAR) test needs, already in registers
  double v[12];       // f64(v2l): row-major 3x4
  double azc;
  const float *incl;  // flipped table
  int H, W, mono;
};

// argmin_h |inc - f64(incl[h])|, first index wins ties (occ_annotate.py:168-173)
__device__ __forceinline__ int nearest_row(double inc, const float *__restrict__ incl, int H, int mono) {
  if (mono == 0) {
    int best = 0;
    double bestd = INFINITY;
    for (int h = 0; h < H; ++h) {
      double d = fabs(__dsub_rn(inc, (double)__ldg(incl + h)));
      if (d < bestd) { bestd = d; best = h; }
    }
    return best;
  }
  // monotone table: k = first index whose entry is on the far side of inc
  int lo = 0, hi = H;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    double v = (double)__ldg(incl + mid);
    bool before = (mono < 0) ? (v > inc) : (v < inc);
    if (before) lo = mid + 1; else hi = mid;
  }
  int best = lo;
  if (lo >= H) {
    best = H - 1;
  } else if (lo > 0) {
    double d0 = fabs(__dsub_rn(inc, (double)__ldg(incl + lo - 1)));
    double d1 = fabs(__dsub_rn(inc, (double)__ldg(incl + lo)));
    best = (d1 < d0) ? lo : lo - 1;   // strict <: the earlier index wins a tie
  }
  while (best > 0 && __ldg(incl + best - 1) == __ldg(incl + best)) --best;   // duplicated entries
  return best;
}

// Returns (row, col, range).  col is the reference's int32 value (may be -1 / needs index wrap).
__device__ __forceinline__ void project_exact(double ex, double ey, double ez, const SensorView &s,
                                              int &row, int &col, double &range) {
  // einsum('bij,bkj->bik') == FMA chain over j, then + translation (occ_annotate.py:164)
  double px = __dadd_rn(__fma_rn(ez, s.v[2], __fma_rn(ey, s.v[1], __dmul_rn(ex, s.v[0]))), s.v[3]);
  double py = __dadd_rn(__fma_rn(ez, s.v[6], __fma_rn(ey, s.v[5], __dmul_rn(ex, s.v[4]))), s.v[7]);
  double pz = __dadd_rn(__fma_rn(ez, s.v[10], __fma_rn(ey, s.v[9], __dmul_rn(ex, s.v[8]))), s.v[11]);
  double xy = __dsqrt_rn(__fma_rn(py, py, __dmul_rn(px, px)));            // :165
  double inc = atan2(pz, xy);                                              // :166
  row = nearest_row(inc, s.incl, s.H, s.mono);
  double az = __dadd_rn(atan2(py, px), s.azc);                             // :176-178
  const double kPi = 3.14159265358979323846;
  const double kTwoPiF32 = 6.2831854820251465;                             // f64(float32(2*pi)), :182,:185
  bool gt = az > kPi, lt = az < -kPi;
  if (gt) az = __dsub_rn(az, kTwoPiF32);
  if (lt) az = __dadd_rn(az, kTwoPiF32);
  double w = (double)s.W;
  double colf = __dsub_rn(__dadd_rn(__dsub_rn(w, 1.0), 0.5),
                          __dmul_rn(__ddiv_rn(__dadd_rn(az, kPi), 6.28318530717958647692), w));  // :187-189
  colf = rint(colf);                                                       // torch.round, half to even (:190)
  colf = fmod(colf, w);                                                    // :191
  col = (int)colf;
  range = __dsqrt_rn(__fma_rn(pz, pz, __fma_rn(py, py, __dmul_rn(px, px))));   // :198
}

}  // namespace occb200
