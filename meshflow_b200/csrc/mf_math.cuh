// Per-element arithmetic of the MeshFlow hot path, shared by every kernel.
//
// Everything that DECIDES an integer (feature->vertex membership, cell membership, 1/32-px source
// coordinates, crop edges, fixed-point blends) is written with explicit round-to-nearest
// multiply/add/divide so that nvcc can never contract it into an FMA: the reference computes these in
// plain IEEE double (NumPy / OpenCV scalar paths), and an FMA changes the last bit.  The functions are
// __host__ __device__ so that the CPU-only test-suite can drive the very same code through
// tests/hostemu (g++ -ffp-contract=off); the shipped library only ever runs them on the GPU.
//
// Reference line numbers are for how4rd/meshflow's meshflowstabilizer.py ("mfs.py").
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define MF_HD __host__ __device__ __forceinline__
#else
#define MF_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define MF_MUL(a, b) __dmul_rn((a), (b))
#define MF_ADD(a, b) __dadd_rn((a), (b))
#define MF_SUB(a, b) __dsub_rn((a), (b))
#define MF_DIV(a, b) __ddiv_rn((a), (b))
#define MF_FMA(a, b, c) __fma_rn((a), (b), (c))
#define MF_SQRT(a) __dsqrt_rn((a))
#define MF_FMUL(a, b) __fmul_rn((a), (b))
#define MF_FADD(a, b) __fadd_rn((a), (b))
#define MF_FSUB(a, b) __fsub_rn((a), (b))
#else
#define MF_MUL(a, b) ((a) * (b))
#define MF_ADD(a, b) ((a) + (b))
#define MF_SUB(a, b) ((a) - (b))
#define MF_DIV(a, b) ((a) / (b))
#define MF_FMA(a, b, c) fma((a), (b), (c))
#define MF_SQRT(a) sqrt((a))
#define MF_FMUL(a, b) ((a) * (b))
#define MF_FADD(a, b) ((a) + (b))
#define MF_FSUB(a, b) ((a) - (b))
#endif

#define MF_INT_MIN (-2147483647 - 1)
#define MF_INT_MAX 2147483647

namespace mf {

// ---------------------------------------------------------------------------------------------
// rounding helpers
// ---------------------------------------------------------------------------------------------
// cvRound / saturate_cast<int>(double): round half to even, clamp, NaN -> INT_MIN.
MF_HD int round_sat(double v) {
  if (!(v == v)) return MF_INT_MIN;
  if (v <= -2147483648.0) return MF_INT_MIN;
  if (v >= 2147483647.0) return MF_INT_MAX;
#if defined(__CUDA_ARCH__)
  return __double2int_rn(v);
#else
  return (int)nearbyint(v);
#endif
}

MF_HD int round_sat_f(float v) {
#if defined(__CUDA_ARCH__)
  const int r = __float2int_rn(v);          // saturates at +-2^31; only NaN needs a hand
  return (v == v) ? r : MF_INT_MIN;
#else
  if (!(v == v)) return MF_INT_MIN;
  if (v <= -2147483648.0f) return MF_INT_MIN;
  if (v >= 2147483520.0f) return MF_INT_MAX;
  return (int)nearbyintf(v);
#endif
}

// ---------------------------------------------------------------------------------------------
// cv2.perspectiveTransform, float64 arithmetic:  w = 1/w (0 when |w| <= DBL_EPSILON), then multiply.
// (mfs.py:325, 420, 1054)
//
// The sums x*m0 + y*m1 + m2 are evaluated the way OpenCV's binary evaluates them on every x86-64 CPU with FMA3
// (perspectiveTransform_ in core/src/matmul.simd.hpp is built per CPU dispatch level -- AVX2, AVX-512 -- with GCC's
// default -ffp-contract=fast):  fma(x, m0, y*m1) + m2  -- y*m1 rounded, x*m0 fused, m2 added last.  Pinned against
// cv2 itself (tests/test_oracle.py) and by the reference's output on videos/video-2, whose first frame pair is almost
// static: its residuals late - H(early) are ~1e-6 px, so a last-bit difference in H(early) shows in the float32
// velocity.  (A CPU without FMA would run the baseline build and round x*m0 separately.)
// ---------------------------------------------------------------------------------------------
MF_HD double persp_sum(double x, double a, double yb, double c) { return MF_ADD(MF_FMA(x, a, yb), c); }

MF_HD void persp(const double* M, double x, double y, double& ox, double& oy) {
  double w = persp_sum(x, M[6], MF_MUL(y, M[7]), M[8]);
  w = (fabs(w) > 2.220446049250313e-16) ? MF_DIV(1.0, w) : 0.0;
  ox = MF_MUL(persp_sum(x, M[0], MF_MUL(y, M[1]), M[2]), w);
  oy = MF_MUL(persp_sum(x, M[3], MF_MUL(y, M[4]), M[5]), w);
}

// ---------------------------------------------------------------------------------------------
// Feature -> vertex membership (mfs.py:426-446).  For a feature at (fx, fy) returns the inclusive
// row window [top, bot]; col_range() gives the inclusive column range of one row in that window.
// ---------------------------------------------------------------------------------------------
struct FeatureCell {
  double frow, fcol;
  int top, bot;
};

MF_HD FeatureCell feature_cell(double fx, double fy, int W, int H, int R, int C, int er) {
  FeatureCell f;
  f.frow = MF_MUL(MF_DIV(fy, (double)H), (double)R);
  f.fcol = MF_MUL(MF_DIV(fx, (double)W), (double)C);
  const double half_rows = MF_DIV((double)er, 2.0);
  double t = ceil(MF_SUB(f.frow, half_rows));
  double b = floor(MF_ADD(f.frow, half_rows));
  // clamp in double first so that wild coordinates cannot overflow the int conversion
  t = t < 0.0 ? 0.0 : (t > (double)(R + 1) ? (double)(R + 1) : t);
  b = b > (double)R ? (double)R : (b < -1.0 ? -1.0 : b);
  f.top = (int)t;
  f.bot = (int)b;
  return f;
}

MF_HD void col_range(const FeatureCell& f, int vr, int C, int er, int ec, int& left, int& right) {
  const double q = MF_DIV(MF_SUB((double)vr, f.frow), (double)er);
  double arg = MF_SUB(0.25, MF_MUL(q, q));
  if (arg < 0.0) arg = 0.0;  // the reference would raise; cannot happen inside [top, bot]
  const double half = MF_MUL((double)ec, MF_SQRT(arg));
  double l = ceil(MF_SUB(f.fcol, half));
  double r = floor(MF_ADD(f.fcol, half));
  l = l < 0.0 ? 0.0 : (l > (double)(C + 1) ? (double)(C + 1) : l);
  r = r > (double)C ? (double)C : (r < -1.0 ? -1.0 : r);
  left = (int)l;
  right = (int)r;
}

// Order-preserving map double -> uint64 (radix select of the per-vertex medians).
MF_HD uint64_t key_of(double v) {
  uint64_t b;
#if defined(__CUDA_ARCH__)
  b = (uint64_t)__double_as_longlong(v);
#else
  union { double d; uint64_t u; } cv; cv.d = v; b = cv.u;
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
MF_HD double value_of(uint64_t k) {
  uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  union { double d; uint64_t u; } cv; cv.u = b; return cv.d;
#endif
}

// median of 9 floats (cv2.medianBlur ksize 3, mfs.py:359-360)
MF_HD void sort2f(float& a, float& b) { float lo = a < b ? a : b; float hi = a < b ? b : a; a = lo; b = hi; }
MF_HD float median9(float* v) {
  sort2f(v[1], v[2]); sort2f(v[4], v[5]); sort2f(v[7], v[8]);
  sort2f(v[0], v[1]); sort2f(v[3], v[4]); sort2f(v[6], v[7]);
  sort2f(v[1], v[2]); sort2f(v[4], v[5]); sort2f(v[7], v[8]);
  sort2f(v[0], v[3]); sort2f(v[5], v[8]); sort2f(v[4], v[7]);
  sort2f(v[3], v[6]); sort2f(v[1], v[4]); sort2f(v[2], v[5]);
  sort2f(v[4], v[7]); sort2f(v[4], v[2]); sort2f(v[6], v[4]);
  sort2f(v[4], v[2]);
  return v[4];
}

// ---------------------------------------------------------------------------------------------
// Adaptive weight lambda_t (mfs.py:786-841): eigenvalue magnitudes of [[a,b,tx],[c,d,ty],[0,0,1]]
// are 1 and those of the 2x2 block.
// ---------------------------------------------------------------------------------------------
MF_HD double adaptive_lambda(const double* Hm, int W, int H, int definition) {
  if (definition == 2) return 100.0;
  if (definition == 3) return 1.0;
  const double a = Hm[0], b = Hm[1], tx = Hm[2], c = Hm[3], d = Hm[4], ty = Hm[5];
  // explicit roundings (no FMA contraction): near a double eigenvalue the discriminant decides the
  // branch, and the oracle's NumPy closed form rounds every product
  const double tr = MF_ADD(a, d);
  const double det = MF_SUB(MF_MUL(a, d), MF_MUL(b, c));
  const double disc = MF_SUB(MF_MUL(tr, tr), MF_MUL(4.0, det));
  double m1, m2;
  if (disc >= 0.0) {
    const double sq = MF_SQRT(disc);
    m1 = fabs(MF_DIV(MF_ADD(tr, sq), 2.0));
    m2 = fabs(MF_DIV(MF_SUB(tr, sq), 2.0));
  } else {
    m1 = m2 = MF_SQRT(fabs(det));
  }
  // sort {1, m1, m2} ascending -> ratio = middle / largest
  double lo = 1.0, mid = m1, hi = m2, t;
  if (lo > mid) { t = lo; lo = mid; mid = t; }
  if (mid > hi) { t = mid; mid = hi; hi = t; }
  if (lo > mid) { t = lo; lo = mid; mid = t; }
  const double ratio = MF_DIV(mid, hi);
  const double qx = MF_DIV(tx, (double)W), qy = MF_DIV(ty, (double)H);
  const double trans = MF_SQRT(MF_ADD(MF_MUL(qx, qx), MF_MUL(qy, qy)));
  const double c1 = MF_ADD(MF_MUL(-1.93, trans), 0.95);
  const double c2 = (definition == 0) ? MF_ADD(MF_MUL(5.83, ratio), 4.88) : MF_SUB(MF_MUL(5.83, ratio), 4.88);
  const double m = c1 < c2 ? c1 : c2;
  return m > 0.0 ? m : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Mesh-cell set-up (mfs.py:1025-1048): exact 4-point homography by 8x8 Gaussian elimination with
// partial pivoting (operation order fixed, identical to oracle/spec.py solve8_partial_pivot).
// ---------------------------------------------------------------------------------------------
MF_HD void homography_4pt(const double* src /*[8] x0,y0,..*/, const double* dst, double* Hout /*[9]*/) {
  double A[8][9];
  for (int i = 0; i < 4; ++i) {
    const double x = src[2 * i], y = src[2 * i + 1], X = dst[2 * i], Y = dst[2 * i + 1];
    double* r0 = A[2 * i];
    double* r1 = A[2 * i + 1];
    r0[0] = x; r0[1] = y; r0[2] = 1.0; r0[3] = 0.0; r0[4] = 0.0; r0[5] = 0.0;
    r0[6] = -MF_MUL(x, X); r0[7] = -MF_MUL(y, X); r0[8] = X;
    r1[0] = 0.0; r1[1] = 0.0; r1[2] = 0.0; r1[3] = x; r1[4] = y; r1[5] = 1.0;
    r1[6] = -MF_MUL(x, Y); r1[7] = -MF_MUL(y, Y); r1[8] = Y;
  }
  for (int k = 0; k < 8; ++k) {
    int piv = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < 8; ++i) {
      const double v = fabs(A[i][k]);
      if (v > best) { best = v; piv = i; }
    }
    if (piv != k) {
      for (int j = 0; j < 9; ++j) { const double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
    }
    for (int i = k + 1; i < 8; ++i) {
      const double f = MF_DIV(A[i][k], A[k][k]);
      for (int j = k; j < 9; ++j) A[i][j] = MF_SUB(A[i][j], MF_MUL(f, A[k][j]));
    }
  }
  double h[8];
  for (int i = 7; i >= 0; --i) {
    double acc = A[i][8];
    for (int j = i + 1; j < 8; ++j) acc = MF_SUB(acc, MF_MUL(A[i][j], h[j]));
    h[i] = MF_DIV(acc, A[i][i]);
  }
  for (int i = 0; i < 8; ++i) Hout[i] = h[i];
  Hout[8] = 1.0;
}

// closed-form inverse (adjugate / determinant), the order of cv::invert for 3x3 (cv2.warpPerspective)
MF_HD void inverse3x3(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double c00 = MF_SUB(MF_MUL(e, i), MF_MUL(f, h));
  const double c01 = MF_SUB(MF_MUL(d, i), MF_MUL(f, g));
  const double c02 = MF_SUB(MF_MUL(d, h), MF_MUL(e, g));
  const double det = MF_ADD(MF_SUB(MF_MUL(a, c00), MF_MUL(b, c01)), MF_MUL(c, c02));
  const double r = (det != 0.0) ? MF_DIV(1.0, det) : 0.0;
  o[0] = MF_MUL(c00, r);
  o[1] = MF_MUL(MF_SUB(MF_MUL(c, h), MF_MUL(b, i)), r);
  o[2] = MF_MUL(MF_SUB(MF_MUL(b, f), MF_MUL(c, e)), r);
  o[3] = MF_MUL(MF_SUB(MF_MUL(f, g), MF_MUL(d, i)), r);
  o[4] = MF_MUL(MF_SUB(MF_MUL(a, i), MF_MUL(c, g)), r);
  o[5] = MF_MUL(MF_SUB(MF_MUL(c, d), MF_MUL(a, f)), r);
  o[6] = MF_MUL(c02, r);
  o[7] = MF_MUL(MF_SUB(MF_MUL(b, g), MF_MUL(a, h)), r);
  o[8] = MF_MUL(MF_SUB(MF_MUL(a, e), MF_MUL(b, d)), r);
}

// One mesh cell of one frame, as the warp kernel consumes it (240 bytes, 16-byte aligned; ordered by
// how often a pixel needs each part: support box, float32 screen, map, float64 membership).
struct alignas(16) Cell {
  int bx0, by0, bx1, by1;       // conservative support box in the output frame (inclusive); bx0 > bx1: empty
  // float32 screening copy of the membership test (decides only when the answer cannot depend on
  // rounding; the float64 test below stays the authority inside the +-eps band around the bounds)
  float fm[9];                  // Mi in box-local coordinates, rounded to float32
  float flo_x, fhi_x, flo_y, fhi_y;   // bounds in source pixels relative to the rest cell's corner
  float feps;                   // screening margin in source pixels; < 0: never screen this cell
  float pad_[2];
  double Hsu[8];                // stabilized -> unstabilized homography (h22 == 1), gives the remap coordinates
  double Mi[9];                 // inverse of the unstabilized -> stabilized homography, decides membership
  int lo_x, hi_x, lo_y, hi_y;   // membership bounds on the 1/32-px source coordinate (inclusive)
  int bounded;                  // 1: the support box is the bounded image of the grown rest rectangle; 0: "anywhere"
  unsigned edge_flags;          // kEdge*: pixels of this cell can satisfy a crop-edge search (mfs.py:1075-1098);
                                // kMapMonotone: the remap denominator keeps its sign over the support box
};

// rest: 4 corners TL,TR,BL,BR of the rest cell (integer valued), stab: the stabilized corners already
// rounded to float32 (cv2.findHomography converts its input to float32).
MF_HD void cell_setup_from_homographies(const double* rest, const double* Hus, const double* Hsu, int W, int H, Cell& out);

MF_HD void cell_setup(const double* rest, const double* stab, int W, int H, Cell& out) {
  double Hus[9], Hsu[9];
  homography_4pt(rest, stab, Hus);
  homography_4pt(stab, rest, Hsu);
  cell_setup_from_homographies(rest, Hus, Hsu, W, H, out);
}

// Everything of cell_setup after the two 4-point solves (the device solves them with eight lanes per system).
MF_HD void cell_setup_from_homographies(const double* rest, const double* Hus, const double* Hsu, int W, int H, Cell& out) {
  for (int i = 0; i < 8; ++i) out.Hsu[i] = Hsu[i];
  inverse3x3(Hus, out.Mi);
  double minx = rest[0], maxx = rest[0], miny = rest[1], maxy = rest[1];
  for (int i = 1; i < 4; ++i) {
    minx = rest[2 * i] < minx ? rest[2 * i] : minx;
    maxx = rest[2 * i] > maxx ? rest[2 * i] : maxx;
    miny = rest[2 * i + 1] < miny ? rest[2 * i + 1] : miny;
    maxy = rest[2 * i + 1] > maxy ? rest[2 * i + 1] : maxy;
  }
  const int L = (int)floor(minx), Rr = (int)ceil(maxx), T = (int)floor(miny), B = (int)ceil(maxy);
  out.lo_x = 32 * L - 31; out.hi_x = 32 * Rr + 31;
  out.lo_y = 32 * T - 31; out.hi_y = 32 * B + 31;
  // Support of the cell in the output frame = image under Hus of the source rectangle grown by one
  // pixel.  While the projective denominator keeps one sign on the four grown corners that image is
  // a bounded convex quad; otherwise fall back to "anywhere".
  const double gx[4] = {(double)L - 1.0, (double)Rr + 1.0, (double)L - 1.0, (double)Rr + 1.0};
  const double gy[4] = {(double)T - 1.0, (double)T - 1.0, (double)B + 1.0, (double)B + 1.0};
  bool pos = true, neg = true, finite = true;
  double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
  for (int i = 0; i < 4; ++i) {
    const double w = gx[i] * Hus[6] + gy[i] * Hus[7] + Hus[8];
    pos = pos && (w > 1e-9);
    neg = neg && (w < -1e-9);
    const double px = (gx[i] * Hus[0] + gy[i] * Hus[1] + Hus[2]) / w;
    const double py = (gx[i] * Hus[3] + gy[i] * Hus[4] + Hus[5]) / w;
    finite = finite && (fabs(px) < 1e9) && (fabs(py) < 1e9);
    x0 = px < x0 ? px : x0; x1 = px > x1 ? px : x1;
    y0 = py < y0 ? py : y0; y1 = py > y1 ? py : y1;
  }
  // float32 screening parameters are filled in below once the support box is known
  for (int i = 0; i < 9; ++i) out.fm[i] = 0.0f;
  out.flo_x = out.flo_y = -31.5f / 32.0f;
  out.fhi_x = (float)(Rr - L) + 31.5f / 32.0f;
  out.fhi_y = (float)(B - T) + 31.5f / 32.0f;
  out.feps = -1.0f;
  out.bounded = ((pos || neg) && finite) ? 1 : 0;
  // Only cells whose rest rectangle reaches within two pixels of the frame border can produce remap
  // coordinates that satisfy a crop-edge search (|m - e| < 1): a member pixel of cell (L..Rr, T..B) has
  // map_x in [L-1, Rr+1], map_y in [T-1, B+1].  (kMapMonotone is added by cell_fast_setup's caller.)
  out.edge_flags = (L <= 2 ? 1u : 0u) | (Rr >= W - 3 ? 2u : 0u) | (T <= 2 ? 4u : 0u) | (B >= H - 3 ? 8u : 0u);
  out.pad_[0] = out.pad_[1] = 0.0f;
  if ((pos || neg) && finite) {
    double fx0 = floor(x0) - 2.0, fx1 = ceil(x1) + 2.0, fy0 = floor(y0) - 2.0, fy1 = ceil(y1) + 2.0;
    fx0 = fx0 < 0.0 ? 0.0 : fx0; fy0 = fy0 < 0.0 ? 0.0 : fy0;
    fx1 = fx1 > (double)(W - 1) ? (double)(W - 1) : fx1;
    fy1 = fy1 > (double)(H - 1) ? (double)(H - 1) : fy1;
    if (fx0 > fx1 || fy0 > fy1) { out.bx0 = 1; out.bx1 = 0; out.by0 = 1; out.by1 = 0; }
    else {
      out.bx0 = (int)fx0; out.bx1 = (int)fx1; out.by0 = (int)fy0; out.by1 = (int)fy1;
      // Screening evaluates Mi in coordinates local to the support box (x' = x - bx0, y' = y - by0)
      // and relative to the rest cell's corner (X' = X - L, Y' = Y - T), so that every float32
      // magnitude is of the order of the cell size, not of the frame size.  It is allowed while the
      // denominator keeps one sign and varies by less than 2x over the box.  Margin: 2^-18 of the
      // magnitudes entering the quotient (>= 16x the worst-case float32 evaluation error).
      const double* M = out.Mi;
      const double d0 = M[6] * fx0 + M[7] * fy0 + M[8];
      const double lm[9] = {M[0] - (double)L * M[6], M[1] - (double)L * M[7],
                            (M[0] * fx0 + M[1] * fy0 + M[2]) - (double)L * d0,
                            M[3] - (double)T * M[6], M[4] - (double)T * M[7],
                            (M[3] * fx0 + M[4] * fy0 + M[5]) - (double)T * d0,
                            M[6], M[7], d0};
      const double bw = fx1 - fx0, bh = fy1 - fy0;
      const double cx[4] = {0.0, bw, 0.0, bw}, cy[4] = {0.0, 0.0, bh, bh};
      double dmin = 1e300, dmax = 0.0;
      bool dpos = true, dneg = true;
      for (int i = 0; i < 4; ++i) {
        const double d = lm[6] * cx[i] + lm[7] * cy[i] + lm[8];
        dpos = dpos && d > 0.0; dneg = dneg && d < 0.0;
        const double a = fabs(d);
        dmin = a < dmin ? a : dmin; dmax = a > dmax ? a : dmax;
      }
      if ((dpos || dneg) && dmin >= 0.5 * dmax && dmin > 0.0) {
        const double span = 2.0 * ((double)(Rr - L) + (double)(B - T)) + 8.0;     // |X'|, |Y'| that matter
        const double mag = (fabs(lm[0]) + fabs(lm[3])) * bw + (fabs(lm[1]) + fabs(lm[4])) * bh + fabs(lm[2]) +
                           fabs(lm[5]) + (fabs(lm[6]) * bw + fabs(lm[7]) * bh + fabs(lm[8])) * span;
        const double eps = mag / dmin * (1.0 / 262144.0);
        if (eps < 0.25) {
          for (int i = 0; i < 9; ++i) out.fm[i] = (float)(lm[i] / d0);   // normalised: denominator ~ 1
          out.feps = (float)eps;
        }
      }
    }
  } else {
    out.bx0 = 0; out.by0 = 0; out.bx1 = W - 1; out.by1 = H - 1;
  }
}

// cv2.warpPerspective(rect mask) != 0 at output pixel (x, y)  (mfs.py:1050-1052)
MF_HD bool cell_inside(const Cell& c, double x, double y) {
  double wd = MF_ADD(MF_ADD(MF_MUL(c.Mi[6], x), MF_MUL(c.Mi[7], y)), c.Mi[8]);
  wd = (wd != 0.0) ? MF_DIV(32.0, wd) : 0.0;
  const int X = round_sat(MF_MUL(MF_ADD(MF_ADD(MF_MUL(c.Mi[0], x), MF_MUL(c.Mi[1], y)), c.Mi[2]), wd));
  const int Y = round_sat(MF_MUL(MF_ADD(MF_ADD(MF_MUL(c.Mi[3], x), MF_MUL(c.Mi[4], y)), c.Mi[5]), wd));
  return X >= c.lo_x && X <= c.hi_x && Y >= c.lo_y && Y <= c.hi_y;
}

// float32 screening of cell_inside: 1 = certainly inside, 0 = certainly outside, -1 = ask cell_inside.
// x is the pixel column relative to the cell's box (px - bx0); (bx, by, bw) are the row parts
// fm1*y' + fm2,  fm4*y' + fm5,  fm7*y' + fm8 with y' = py - by0, hoisted by the caller.
MF_HD int cell_screen(const Cell& c, float x, float bx, float by, float bw) {
  if (c.feps < 0.0f) return -1;
#if defined(__CUDA_ARCH__)
  const float wd = fmaf(c.fm[6], x, bw);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(wd));
  const float X = fmaf(c.fm[0], x, bx) * r;
  const float Y = fmaf(c.fm[3], x, by) * r;
#else
  const float wd = c.fm[6] * x + bw;
  const float r = 1.0f / wd;
  const float X = (c.fm[0] * x + bx) * r;
  const float Y = (c.fm[3] * x + by) * r;
#endif
  const float e = c.feps;
  const bool in = X > c.flo_x + e && X < c.fhi_x - e && Y > c.flo_y + e && Y < c.fhi_y - e;
  const bool out = X < c.flo_x - e || X > c.fhi_x + e || Y < c.flo_y - e || Y > c.fhi_y + e;
  return in ? 1 : (out ? 0 : -1);   // NaN compares false both ways -> -1
}

// Screening with the outcome split by side, for the end-point argument of the warp kernel: the
// region "certainly inside" is convex in the output frame (pre-image of a rectangle under a
// projective map whose denominator keeps its sign), and so is each of the four "certainly beyond
// this bound" regions; two end points in the same region put the whole segment in it.
//   bit 0: certainly inside.  bits 1..4: certainly left of lo_x / right of hi_x / above lo_y / below hi_y.
// Arguments are the cell's float32 parameters as scalars so that callers can keep them in registers.
MF_HD unsigned screen_sides(float m0, float m3, float m6, float bx, float by, float bw, float x,
                            float lo_x, float hi_x, float lo_y, float hi_y, float e) {
#if defined(__CUDA_ARCH__)
  const float wd = fmaf(m6, x, bw);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(wd));
  const float X = fmaf(m0, x, bx) * r;
  const float Y = fmaf(m3, x, by) * r;
#else
  const float wd = m6 * x + bw;
  const float r = 1.0f / wd;
  const float X = (m0 * x + bx) * r;
  const float Y = (m3 * x + by) * r;
#endif
  unsigned m = 0u;
  if (X > lo_x + e && X < hi_x - e && Y > lo_y + e && Y < hi_y - e) m |= 1u;
  if (X < lo_x - e) m |= 2u;
  if (X > hi_x + e) m |= 4u;
  if (Y < lo_y - e) m |= 8u;
  if (Y > hi_y + e) m |= 16u;
  return m;
}

MF_HD unsigned cell_screen_sides(const Cell& c, float x, float bx, float by, float bw) {
  return screen_sides(c.fm[0], c.fm[3], c.fm[6], bx, by, bw, x, c.flo_x, c.fhi_x, c.flo_y, c.fhi_y, c.feps);
}

// float32 remap coordinates of output pixel (x, y) through the cell (mfs.py:1054)
MF_HD void cell_map(const Cell& c, double x, double y, float& mx, float& my) {
  double w = persp_sum(x, c.Hsu[6], MF_MUL(y, c.Hsu[7]), 1.0);
  w = (fabs(w) > 2.220446049250313e-16) ? MF_DIV(1.0, w) : 0.0;
  mx = (float)MF_MUL(persp_sum(x, c.Hsu[0], MF_MUL(y, c.Hsu[1]), c.Hsu[2]), w);
  my = (float)MF_MUL(persp_sum(x, c.Hsu[3], MF_MUL(y, c.Hsu[4]), c.Hsu[5]), w);
}

// cell_map with the row products y*Hsu[1], y*Hsu[4], y*Hsu[7] supplied by the caller: every rounding
// is the one cell_map performs, so the result is bit-identical; 1/w is the correctly rounded
// reciprocal either way.
// Correctly rounded 1/w for a finite, normal w well inside the exponent range (the remap denominator
// is ~1): hardware seed (2^-23), two Newton steps (-> within an ulp), then Markstein's final
// correction y + y*(1 - w*y), which rounds correctly (the only exception, an all-ones significand,
// has probability 2^-52 and would cost one ulp).  Same value as __drcp_rn / IEEE 1.0/w without the
// special-case handling; checked against __drcp_rn on the device by mf_debug_rcp_mismatches.
#if defined(__CUDACC__)
__device__ __forceinline__ double rcp_rn_normal(double w) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
  double e = fma(-w, y, 1.0); y = fma(y, e, y);
  e = fma(-w, y, 1.0); y = fma(y, e, y);
  e = fma(-w, y, 1.0); y = fma(y, e, y);
  return y;
}
#endif

MF_HD void map_row(double h0, double h2, double h3, double h5, double h6, double x, double yh1, double yh4,
                   double yh7, float& mx, float& my) {
  double w = persp_sum(x, h6, yh7, 1.0);
#if defined(__CUDA_ARCH__)
  const double aw = fabs(w);
  w = (aw > 1e-100 && aw < 1e100) ? rcp_rn_normal(w) : ((aw > 2.220446049250313e-16) ? __drcp_rn(w) : 0.0);
#else
  w = (fabs(w) > 2.220446049250313e-16) ? (1.0 / w) : 0.0;
#endif
  mx = (float)MF_MUL(persp_sum(x, h0, yh1, h2), w);
  my = (float)MF_MUL(persp_sum(x, h3, yh4, h5), w);
}

MF_HD void cell_map_row(const Cell& c, double x, double yh1, double yh4, double yh7, float& mx, float& my) {
  double w = persp_sum(x, c.Hsu[6], yh7, 1.0);
#if defined(__CUDA_ARCH__)
  w = (fabs(w) > 2.220446049250313e-16) ? __drcp_rn(w) : 0.0;
#else
  w = (fabs(w) > 2.220446049250313e-16) ? (1.0 / w) : 0.0;
#endif
  mx = (float)MF_MUL(persp_sum(x, c.Hsu[0], yh1, c.Hsu[2]), w);
  my = (float)MF_MUL(persp_sum(x, c.Hsu[3], yh4, c.Hsu[5]), w);
}

// ---------------------------------------------------------------------------------------------
// cv2.remap 8UC3 INTER_LINEAR BORDER_CONSTANT (mfs.py:1063-1069)
// ---------------------------------------------------------------------------------------------
MF_HD void remap_coords(float mx, float my, int& ix, int& iy, int& ax, int& ay) {
  const int sx = round_sat_f(MF_FMUL(mx, 32.0f));
  const int sy = round_sat_f(MF_FMUL(my, 32.0f));
  ix = sx >> 5; iy = sy >> 5; ax = sx & 31; ay = sy & 31;
}

// remap_coords for map values that are known not to be NaN (every value a cell produces is finite:
// a cell with a non-finite homography is never "inside"); saves the NaN select of round_sat_f.
MF_HD void remap_coords_finite(float mx, float my, int& ix, int& iy, int& ax, int& ay) {
#if defined(__CUDA_ARCH__)
  const int sx = __float2int_rn(MF_FMUL(mx, 32.0f));
  const int sy = __float2int_rn(MF_FMUL(my, 32.0f));
  ix = sx >> 5; iy = sy >> 5; ax = sx & 31; ay = sy & 31;
#else
  remap_coords(mx, my, ix, iy, ax, ay);
#endif
}

MF_HD int blend4(int p00, int p01, int p10, int p11, int ax, int ay) {
  return (p00 * (32 - ax) * (32 - ay) + p01 * ax * (32 - ay) + p10 * (32 - ax) * ay + p11 * ax * ay + 512) >> 10;
}

// Slow-but-general pixel fetch with the constant border.
MF_HD void remap_pixel(const uint8_t* src, int W, int H, int ix, int iy, int ax, int ay,
                       int bb, int bg, int br, uint8_t* out3) {
  const int bord[3] = {bb, bg, br};
  const bool x0ok = ix >= 0 && ix < W, x1ok = ix + 1 >= 0 && ix + 1 < W;
  const bool y0ok = iy >= 0 && iy < H, y1ok = iy + 1 >= 0 && iy + 1 < H;
  const uint8_t* r0 = src + (size_t)(y0ok ? iy : 0) * W * 3;
  const uint8_t* r1 = src + (size_t)(y1ok ? iy + 1 : 0) * W * 3;
  const int xa = x0ok ? ix : 0, xb = x1ok ? ix + 1 : 0;
  for (int ch = 0; ch < 3; ++ch) {
    const int p00 = (x0ok && y0ok) ? r0[3 * xa + ch] : bord[ch];
    const int p01 = (x1ok && y0ok) ? r0[3 * xb + ch] : bord[ch];
    const int p10 = (x0ok && y1ok) ? r1[3 * xa + ch] : bord[ch];
    const int p11 = (x1ok && y1ok) ? r1[3 * xb + ch] : bord[ch];
    out3[ch] = (uint8_t)blend4(p00, p01, p10, p11, ax, ay);
  }
}

// ---------------------------------------------------------------------------------------------
// cv2.resize 8UC3 INTER_LINEAR, 11-bit fixed point (mfs.py:1150-1155)
// One axis: source index pair and the two integer weights of destination index d.
// clamp_frac = true for the x axis (fraction zeroed at the borders), false for y (rows clipped only).
// ---------------------------------------------------------------------------------------------
MF_HD void resize_coef(int d, int src_len, int dst_len, bool is_x, int& i0, int& i1, int& w0, int& w1) {
  const double scale = MF_DIV((double)src_len, (double)dst_len);
  float f = (float)MF_SUB(MF_MUL(MF_ADD((double)d, 0.5), scale), 0.5);
  int s = (int)floorf(f);
  f = MF_FSUB(f, (float)s);
  if (is_x) {
    if (s < 0) { s = 0; f = 0.0f; }
    if (s >= src_len - 1) { s = src_len - 1; f = 0.0f; }
    i0 = s;
    i1 = (s + 1 < src_len) ? s + 1 : src_len - 1;
  } else {
    i0 = s < 0 ? 0 : (s > src_len - 1 ? src_len - 1 : s);
    i1 = s + 1 < 0 ? 0 : (s + 1 > src_len - 1 ? src_len - 1 : s + 1);
  }
  w1 = round_sat_f(MF_FMUL(f, 2048.0f));
  w0 = round_sat_f(MF_FMUL(MF_FSUB(1.0f, f), 2048.0f));
}

MF_HD int resize_blend(int p00, int p01, int p10, int p11, int a0, int a1, int b0, int b1) {
  const int s0 = a0 * p00 + a1 * p01;
  const int s1 = a0 * p10 + a1 * p11;
  int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}


// ---------------------------------------------------------------------------------------------
// Fast path of the warp (warp_fast.cuh): exact row spans per cell + float32 remap coordinates with a
// proven rounding band.
//
// (1) Row spans.  For a cell whose membership denominator keeps one sign over its support box, the
//     test cell_inside() is, in real arithmetic, four half-planes alpha*x + beta*y + gamma >= 0 in
//     output-pixel coordinates: X >= lo_x  <=>  32*NX - (lo_x - 1/2)*D >= 0 (D > 0), etc.  On one
//     output row the member pixels are therefore one interval [a, b].  span_of_row() finds it from
//     the half-planes and settles the (at most one per half-plane) pixel that lies within the
//     rounding noise of a boundary with the exact test, so the interval equals cell_inside() pixel
//     for pixel.  "noise" bounds |reference float64 evaluation - real arithmetic| with a 32x margin.
// (2) Float32 remap coordinates.  U = 32*(map_x - X0), V = 32*(map_y - Y0) evaluated in float32 in
//     box-centred coordinates differ from the real-arithmetic value by less than eps (derivation in
//     DESIGN.md section 4); when U is farther than eps from a rounding boundary (k + 1/2), rint(U) is
//     the reference's rint(32 * float32(map_x)) - 32*L0 exactly: rounding to float32 is monotone and
//     every tie (k + 1/2)/32 is a float32 (|map| < 2^17), so float32(map_x) stays on its side of the
//     tie -- unless map_x is within half a float32 ulp of it, where it can land ON the tie and
//     round-half-even decides; that radius (16 ulp of the largest |map| of the cell, in 1/32-px
//     units) is part of the band.  Pixels inside the band take the float64 path.
// ---------------------------------------------------------------------------------------------
struct alignas(16) CellFast {
  float a[9];          // U = (a0 x' + a1 y' + a2) / (a6 x' + a7 y' + a8), V = (a3 x' + a4 y' + a5) / (same); a8 == 1
  float thr_u;         // a pixel is safe when |U - rint(U)| <= thr_u and |V - rint(V)| <= thr_v; thr_u < 0: no fast path
  int bx0, by0;        // x' = px - bx0, y' = py - by0 (centre of the support box)
  int base_x, base_y;  // 32 * centre of the rest cell: source 1/32-px coordinate = base + rint(U)
  unsigned flags;      // kEdge*: pixels of this cell can satisfy a crop-edge search (mfs.py:1075-1098)
  float thr_v;
};

struct alignas(16) CellSpan {
  double al[4], be[4], ga[4];   // half-planes, already multiplied by the sign of the denominator
  double ial[4], mrg[4];        // -1/al and the width (pixels) of the uncertain zone around the crossing; ial = 0: degenerate
  double noise;                 // |evaluated half-plane - truth| bound
  int regular;                  // 1: spans are valid; 0: resolve this cell pixel by pixel
  int pad_;
};

static constexpr unsigned kEdgeLeft = 1u, kEdgeRight = 2u, kEdgeTop = 4u, kEdgeBottom = 8u;
static constexpr unsigned kEdgeAny = 15u;
static constexpr unsigned kMapMonotone = 16u;    // Cell::edge_flags only: the cell has a float32 map form (thr_u >= 0)
static constexpr unsigned kSegNone = 0xffffu;    // no cell covers the segment: default map, border colour
static constexpr unsigned kSegIrregular = 0xfffeu;  // resolve every pixel of this row segment exactly
static constexpr unsigned kSegStraddle = 0xfffdu;   // lane-owner table only: a segment starts inside the group of four
static constexpr unsigned kSegSentinel = 0xffffffffu;
static constexpr int kSegMax = 16;

// rest rectangle bounds L, Rr, T, B as cell_setup computes them (floor/ceil of the rest corners)
MF_HD void cell_fast_setup(const Cell& c, int L, int Rr, int T, int B, int W, int H, CellFast& cf, CellSpan& sp) {
  for (int i = 0; i < 9; ++i) cf.a[i] = 0.0f;
  cf.thr_u = cf.thr_v = -1.0f; cf.bx0 = c.bx0; cf.by0 = c.by0;
  cf.base_x = cf.base_y = 0;
  cf.flags = (L <= 2 ? kEdgeLeft : 0u) | (Rr >= W - 3 ? kEdgeRight : 0u) | (T <= 2 ? kEdgeTop : 0u) |
             (B >= H - 3 ? kEdgeBottom : 0u);
  for (int i = 0; i < 4; ++i) { sp.al[i] = sp.be[i] = sp.ga[i] = sp.ial[i] = sp.mrg[i] = 0.0; }
  sp.noise = 0.0; sp.regular = 0; sp.pad_ = 0;
  if (!c.bounded || c.bx0 > c.bx1) return;
  const double fx0 = (double)c.bx0, fx1 = (double)c.bx1, fy0 = (double)c.by0, fy1 = (double)c.by1;
  const double cx[4] = {fx0, fx1, fx0, fx1}, cy[4] = {fy0, fy0, fy1, fy1};
  // ---- spans: membership half-planes ----
  {
    const double* M = c.Mi;
    bool dpos = true, dneg = true, fin = true;
    double dmin = 1e300, dmax = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double d = M[6] * cx[i] + M[7] * cy[i] + M[8];
      dpos = dpos && d > 0.0; dneg = dneg && d < 0.0;
      const double a = fabs(d);
      dmin = a < dmin ? a : dmin; dmax = a > dmax ? a : dmax;
    }
    for (int i = 0; i < 9; ++i) fin = fin && (fabs(M[i]) < 1e100);
    if ((dpos || dneg) && fin && dmin > 1e-100 && dmax < 1e100) {
      const double sg = dpos ? 1.0 : -1.0;
      const double kx0 = (double)c.lo_x - 0.5, kx1 = (double)c.hi_x + 0.5;
      const double ky0 = (double)c.lo_y - 0.5, ky1 = (double)c.hi_y + 0.5;
      sp.al[0] = sg * (32.0 * M[0] - kx0 * M[6]); sp.be[0] = sg * (32.0 * M[1] - kx0 * M[7]); sp.ga[0] = sg * (32.0 * M[2] - kx0 * M[8]);
      sp.al[1] = sg * (kx1 * M[6] - 32.0 * M[0]); sp.be[1] = sg * (kx1 * M[7] - 32.0 * M[1]); sp.ga[1] = sg * (kx1 * M[8] - 32.0 * M[2]);
      sp.al[2] = sg * (32.0 * M[3] - ky0 * M[6]); sp.be[2] = sg * (32.0 * M[4] - ky0 * M[7]); sp.ga[2] = sg * (32.0 * M[5] - ky0 * M[8]);
      sp.al[3] = sg * (ky1 * M[6] - 32.0 * M[3]); sp.be[3] = sg * (ky1 * M[7] - 32.0 * M[4]); sp.ga[3] = sg * (ky1 * M[8] - 32.0 * M[5]);
      const double w = (double)W, h = (double)H;
      const double magD = fabs(M[6]) * w + fabs(M[7]) * h + fabs(M[8]);
      const double magX = fabs(M[0]) * w + fabs(M[1]) * h + fabs(M[2]);
      const double magY = fabs(M[3]) * w + fabs(M[4]) * h + fabs(M[5]);
      const double magN = magX > magY ? magX : magY;
      const double kmax = 32.0 * (w + h) + 64.0;          // |lo|, |hi| + 1/2 never exceed this
      const double u48 = 3.5527136788005009e-15;          // 2^-48: 32x the rounding unit times the ~6 roundings involved
      const double ref = u48 * (32.0 * magN / dmin + kmax * (magD / dmin + 1.0));   // reference X, Y vs truth (1/32-px units)
      sp.noise = 2.0 * (ref * dmax + u48 * (32.0 * magN + kmax * magD));
      for (int i = 0; i < 4; ++i) {
        const double aal = fabs(sp.al[i]);
        if (aal >= 4.0 * sp.noise) { sp.ial[i] = -1.0 / sp.al[i]; sp.mrg[i] = sp.noise / aal + 1e-7; }   // mrg < 0.26
      }
      sp.regular = 1;
    }
  }
  // ---- float32 remap coordinates, centred on the box / on the rest cell (halves every magnitude) ----
  {
    const double* h = c.Hsu;
    const int ocx = (c.bx0 + c.bx1) >> 1, ocy = (c.by0 + c.by1) >> 1;      // output-space origin
    const int scx = (L + Rr) >> 1, scy = (T + B) >> 1;                     // source-space origin
    cf.bx0 = ocx; cf.by0 = ocy; cf.base_x = 32 * scx; cf.base_y = 32 * scy;
    const double ox = (double)ocx, oy = (double)ocy;
    const double d0 = h[6] * ox + h[7] * oy + 1.0;
    bool dpos = true, dneg = true;
    double dmin = 1e300, dmax = 0.0;
    for (int i = 0; i < 4; ++i) {
      const double d = h[6] * cx[i] + h[7] * cy[i] + 1.0;
      dpos = dpos && d > 0.0; dneg = dneg && d < 0.0;
      const double a = fabs(d);
      dmin = a < dmin ? a : dmin; dmax = a > dmax ? a : dmax;
    }
    if (!((dpos || dneg) && dmin >= 0.5 * dmax && dmin > 1e-100 && dmax < 1e100)) return;
    const double l0 = (double)scx, t0 = (double)scy;
    double l[9] = {32.0 * (h[0] - l0 * h[6]), 32.0 * (h[1] - l0 * h[7]), 32.0 * ((h[0] * ox + h[1] * oy + h[2]) - l0 * d0),
                   32.0 * (h[3] - t0 * h[6]), 32.0 * (h[4] - t0 * h[7]), 32.0 * ((h[3] * ox + h[4] * oy + h[5]) - t0 * d0),
                   h[6], h[7], d0};
    for (int i = 0; i < 9; ++i) l[i] /= d0;
    const double bw = (fx1 - fx0) * 0.5 + 1.0, bh = (fy1 - fy0) * 0.5 + 1.0;   // |x'|, |y'| over the box
    const double su = fabs(l[0]) * bw + fabs(l[1]) * bh + fabs(l[2]);
    const double sv = fabs(l[3]) * bw + fabs(l[4]) * bh + fabs(l[5]);
    const double sd = fabs(l[6]) * bw + fabs(l[7]) * bh + 1.0;
    // member pixels map to [L - 1, Rr + 1] x [T - 1, B + 1]: |U|, |V| <= 32 * (half extent + 2)
    const double umax = 32.0 * (double)(((Rr - L > B - T ? Rr - L : B - T) + 1) / 2 + 2);
    const double dn = dmin / fabs(d0);
    const double u24 = 5.9604644775390625e-08;             // 2^-24
    // |float32 value - truth| <= u*(3.01*S + U_max*(3.01*S_d + 3*d_n))/d_n + O(u^2)  (numerator and denominator: rounded
    // coefficients + two FMAs each; rcp.approx: 1 ulp; one product), d_n <= 1 <= S_d
    const double eps = u24 * (3.1 * (su > sv ? su : sv) + 6.2 * umax * sd) / dn + 1e-5;
    // half a float32 ulp of the absolute coordinate, in 1/32-px units: 16 * 2^(e-23) for |coordinate| < 2^(e+1)
    double tie_u = 16.0 * 1.1920928955078125e-07, tie_v = tie_u;
    const int al = L - 1 < 0 ? 1 - L : L - 1, at = T - 1 < 0 ? 1 - T : T - 1;
    const double xm = (double)(al > Rr + 1 ? al : Rr + 1) + 1.0;
    const double ym = (double)(at > B + 1 ? at : B + 1) + 1.0;
    for (double p = 2.0; p <= xm; p *= 2.0) tie_u *= 2.0;
    for (double p = 2.0; p <= ym; p *= 2.0) tie_v *= 2.0;
    if (!(eps + tie_u < 0.2) || !(eps + tie_v < 0.2) || !(su < 1e6) || !(sv < 1e6) || !(xm < 1e5) || !(ym < 1e5)) return;
    for (int i = 0; i < 9; ++i) cf.a[i] = (float)l[i];
    cf.thr_u = (float)(0.5 - 1.001 * (eps + tie_u));
    cf.thr_v = (float)(0.5 - 1.001 * (eps + tie_v));
  }
}

// Member pixels of cell c on output row y within columns [xlo, xhi] (already clipped to the cell's
// box).  Returns 0 and the inclusive interval [a, b]; 1 when no pixel is a member; 2 when the row
// cannot be described by an interval with certainty (caller resolves it pixel by pixel).
MF_HD int span_of_row(const Cell& c, const CellSpan& sp, int y, int xlo, int xhi, int& a, int& b) {
  const double yd = (double)y;
  double lo = (double)xlo, hi = (double)xhi;
  for (int i = 0; i < 4; ++i) {
    const double al = sp.al[i];
    const double rowc = sp.be[i] * yd + sp.ga[i];
    if (sp.ial[i] == 0.0) {
      // (almost) no dependence on x over any frame width the library accepts: decide the row as a whole
      const double c0 = al * (double)xlo + rowc, c1 = al * (double)xhi + rowc;
      const double cmin = c0 < c1 ? c0 : c1, cmax = c0 < c1 ? c1 : c0;
      if (cmin > sp.noise) continue;
      if (cmax < -sp.noise) return 1;
      return 2;
    }
    // crossing of the half-plane with the row; the product differs from the quotient -rowc/al by an ulp,
    // far inside the margin
    const double t = rowc * sp.ial[i];
    const double m = sp.mrg[i];
    if (!(t == t) || fabs(t) > 1e15) return 2;
    if (al > 0.0) {                                        // pixels x >= t
      double first = floor(t + m) + 1.0;                   // smallest integer certainly on the member side
      const double p = first - 1.0;
      if (p >= t - m && p >= lo && p <= hi && cell_inside(c, p, yd)) first = p;
      if (first > lo) lo = first;
    } else {                                               // pixels x <= t
      double last = ceil(t - m) - 1.0;
      const double p = last + 1.0;
      if (p <= t + m && p >= lo && p <= hi && cell_inside(c, p, yd)) last = p;
      if (last < hi) hi = last;
    }
    if (lo > hi) return 1;
  }
  a = (int)lo; b = (int)hi;
  return 0;
}

// Row segments of one 128-pixel tile row: which cell owns each pixel ("the last cell written wins",
// mfs.py:1060-1061).  Candidates are visited by descending id; each takes what is still uncovered
// of its span.  seg[i] = (first x << 16) | cell id, ascending x, seg[0] starts at x0; unused entries
// are kSegSentinel.  finish() returns the number of segments, or -1 when more than CAP are needed.
//
// The state lives in registers (round 2 kept interval and segment lists with run-time indices, i.e. in local
// memory, and the kernel waited on them): the uncovered pixels are a 128-bit mask, a candidate's share is
// mask & span, its runs start where a set bit follows a clear one, and a new segment is put in its place
// of the ascending list by one min / max step per slot -- every index is a compile-time constant.
MF_HD uint32_t shl_sat(uint32_t v, int s) {        // s >= 0; 32 and more shift everything out
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));   // PTX clamps the amount to the register width
  return r;
#else
  return s >= 32 ? 0u : (v << s);
#endif
}
MF_HD uint32_t shr_sat(uint32_t v, int s) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
  return r;
#else
  return s >= 32 ? 0u : (v >> s);
#endif
}
MF_HD int lowest_set_bit(uint32_t v) {             // v != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

template <int CAP>
struct SegBuilder {
  static_assert(CAP >= 1 && CAP <= kSegMax, "segment capacity");
  uint32_t U[4];                 // uncovered pixels: bit i of word w = pixel x0 + 32 w + i
  unsigned seg[CAP];             // ascending, kSegSentinel behind the ns entries
  int ns, x0;
  bool overflow;
  MF_HD void begin(int x0_, int x1) {
    x0 = x0_;
    const int n = x1 - x0 + 1;                               // 1 .. 128
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int cnt = n - 32 * w;                            // pixels of this word and the following ones
      U[w] = cnt <= 0 ? 0u : shr_sat(0xffffffffu, cnt >= 32 ? 0 : 32 - cnt);
    }
#pragma unroll
    for (int i = 0; i < CAP; ++i) seg[i] = kSegSentinel;
    ns = 0; overflow = false;
  }
  MF_HD void emit(int x, unsigned id) {
    if (ns >= CAP) { overflow = true; return; }
    unsigned v = ((unsigned)x << 16) | id;
#pragma unroll
    for (int i = 0; i < CAP; ++i) {                          // sorted insertion: the list stays ascending
      const unsigned lo = seg[i] < v ? seg[i] : v, hi = seg[i] < v ? v : seg[i];
      seg[i] = lo; v = hi;
    }
    ++ns;
  }
  MF_HD void emit_runs(const uint32_t (&T)[4], unsigned id) {
    uint32_t carry = 0u;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      uint32_t S = T[w] & ~((T[w] << 1) | carry);           // first pixel of every run
      carry = T[w] >> 31;
      while (S != 0u) {
        emit(x0 + 32 * w + lowest_set_bit(S), id);
        S &= S - 1u;
      }
    }
  }
  // the candidate `id` takes what is still uncovered of [a, b]  (x0 <= a <= b <= x1)
  MF_HD void cover(int a, int b, unsigned id, int /*cap*/ = CAP) {
    const int la = a - x0, lb = b - x0;
    uint32_t T[4], any = 0u;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int lo = la - 32 * w, hs = 31 - lb + 32 * w;
      const uint32_t M = shl_sat(0xffffffffu, lo > 0 ? lo : 0) & shr_sat(0xffffffffu, hs > 0 ? hs : 0);
      T[w] = U[w] & M;
      U[w] &= ~M;
      any |= T[w];
    }
    if (any != 0u) emit_runs(T, id);
  }
  MF_HD bool done() const { return (U[0] | U[1] | U[2] | U[3]) == 0u; }
  MF_HD int finish(int /*cap*/ = CAP) {
    if (!done()) emit_runs(U, kSegNone);                     // pixels no candidate covers: default map
    return overflow ? -1 : ns;
  }
};

// Owner of pixel x in a sorted segment list (slow path; the fast path does this per 4-pixel group).
MF_HD unsigned seg_owner(const unsigned* seg, int cap, int x) {
  const unsigned key = ((unsigned)x << 16) | 0xffffu;
  unsigned cur = seg[0];
  for (int i = 1; i < cap; ++i) {
    if (seg[i] == kSegSentinel) break;
    if (seg[i] <= key) cur = seg[i];
  }
  return cur & 0xffffu;
}

static constexpr float kRoundMagic = 12582912.0f;          // 1.5 * 2^23: adding it rounds to an integer (RNE)
static constexpr unsigned kRoundMagicBits = 0x4b400000u;

// Float32 remap coordinate of one pixel: nU = rint(U) + kRoundMagicBits as raw bits; returns whether
// the pixel is outside the rounding band (safe).  bx, by, bw: row parts a1*y'+a2, a4*y'+a5, a7*y'+a8.
MF_HD bool fast_coords(float a0, float a3, float a6, float bx, float by, float bw, float x, float thr_u, float thr_v,
                       unsigned& nU, unsigned& nV) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(a6, x, bw)));
  const float U = __fmul_rn(fmaf(a0, x, bx), r), V = __fmul_rn(fmaf(a3, x, by), r);
  const float tU = __fadd_rn(U, kRoundMagic), tV = __fadd_rn(V, kRoundMagic);
  const float dU = __fsub_rn(U, __fsub_rn(tU, kRoundMagic)), dV = __fsub_rn(V, __fsub_rn(tV, kRoundMagic));
  nU = __float_as_uint(tU); nV = __float_as_uint(tV);
#else
  const float r = 1.0f / fmaf(a6, x, bw);
  const float U = fmaf(a0, x, bx) * r, V = fmaf(a3, x, by) * r;
  volatile float tU = U + kRoundMagic, tV = V + kRoundMagic;
  const float dU = U - (tU - kRoundMagic), dV = V - (tV - kRoundMagic);
  union { float f; unsigned u; } cu, cv; cu.f = tU; cv.f = tV;
  nU = cu.u; nV = cv.u;
#endif
  return fabsf(dU) <= thr_u && fabsf(dV) <= thr_v;         // NaN compares false: not safe
}


// One group of four adjacent output pixels (px0..px0+3, py) owned by one cell: raw rounded coordinates
// nu[j], nv[j] (rint + kRoundMagicBits) and the mask of pixels inside the rounding band.
MF_HD unsigned fast_group_coords(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
                                 float a8, float thr_u, float thr_v, int cbx0, int cby0, int px0, int py, unsigned* nu,
                                 unsigned* nv) {
  const float fy = (float)(py - cby0);
  const float bx = fmaf(a1, fy, a2), by = fmaf(a4, fy, a5), bw = fmaf(a7, fy, a8);
  const float fx = (float)(px0 - cbx0);
  unsigned bad = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 4; ++j)
    if (!fast_coords(a0, a3, a6, bx, by, bw, fx + (float)j, thr_u, thr_v, nu[j], nv[j])) bad |= 1u << j;
  return bad;
}

MF_HD unsigned umin4(const unsigned* v) { unsigned a = v[0] < v[1] ? v[0] : v[1], b = v[2] < v[3] ? v[2] : v[3]; return a < b ? a : b; }
MF_HD unsigned umax4(const unsigned* v) { unsigned a = v[0] > v[1] ? v[0] : v[1], b = v[2] > v[3] ? v[2] : v[3]; return a > b ? a : b; }

// What to do with the group: returns the mask of pixels for the slow path (15 = the whole group) and
// whether the group takes the shared-window gather (`fast`).  The gather needs adjacent footprints
// (pixel j reads source columns ix0+j, ix0+j+1 of rows iy0, iy0+1), every tap -- and the 5-word row
// reads -- inside the frame, and no pixel that could satisfy a crop-edge search (those are decided on
// the exact float32 map; a pixel in the rounding band is off by at most one unit, hence the +-2).
MF_HD unsigned fast_group_plan(const unsigned* nu, const unsigned* nv, unsigned bad, int base_x, int base_y,
                               unsigned flags, int W, int H, bool bounds_only, int& ix0, int& iy0, bool& fast, bool& edge) {
  const unsigned bu = nu[0] & ~31u, bv = nv[0] & ~31u;
  const unsigned spread = (nu[1] - bu - 32u) | (nu[2] - bu - 64u) | (nu[3] - bu - 96u) | (nv[1] - bv) | (nv[2] - bv) |
                          (nv[3] - bv);
  ix0 = (int)(bu - kRoundMagicBits + (unsigned)base_x) >> 5;
  iy0 = (int)(bv - kRoundMagicBits + (unsigned)base_y) >> 5;
  edge = false;
  if (flags != 0u) {
    const int nx_lo = (int)(umin4(nu) - kRoundMagicBits) + base_x, nx_hi = (int)(umax4(nu) - kRoundMagicBits) + base_x;
    const int ny_lo = (int)(umin4(nv) - kRoundMagicBits) + base_y, ny_hi = (int)(umax4(nv) - kRoundMagicBits) + base_y;
    edge = ((flags & kEdgeLeft) && nx_lo < 32 + 2) || ((flags & kEdgeRight) && nx_hi > 32 * (W - 2) - 2) ||
           ((flags & kEdgeTop) && ny_lo < 32 + 2) || ((flags & kEdgeBottom) && ny_hi > 32 * (H - 2) - 2);
  }
  fast = false;
  if (bounds_only) return edge ? 15u : 0u;
  const bool ok = !edge && spread < 32u && ix0 >= 0 && ix0 <= W - 8 && iy0 >= 0 && iy0 <= H - 2;
  fast = ok;
  return ok ? bad : 15u;
}

// Owner of the group [px0, px0+3] in a sorted segment list and whether a segment starts inside it.
MF_HD unsigned seg_group_owner(const unsigned* seg, int cap, int px0, bool& straddle) {
  const unsigned key = ((unsigned)px0 << 16) | 0xffffu;
  unsigned cur = seg[0];
  straddle = false;
  for (int i = 1; i < cap; ++i) {
    const unsigned v = seg[i];
    if (v == kSegSentinel) break;
    if (v <= key) cur = v;
    straddle = straddle || ((v - key - 1u) < 0x30000u);
  }
  return cur & 0xffffu;
}


// One pixel on its own (a pixel whose GROUP failed a shape condition): float32 coordinate of the pixel through its
// own cell; true + the 1/32-px source coordinate (sx, sy) when the pixel is outside the rounding band and cannot
// satisfy a crop-edge search, false when it needs the float64 sequence after all.
MF_HD bool medium_coords(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8,
                         float thr_u, float thr_v, int cbx0, int cby0, int base_x, int base_y, unsigned flags, int px,
                         int py, int W, int H, int& sx, int& sy) {
  if (!(thr_u >= 0.0f)) return false;
  const float fy = (float)(py - cby0), fx = (float)(px - cbx0);
  unsigned nu, nv;
  if (!fast_coords(a0, a3, a6, fmaf(a1, fy, a2), fmaf(a4, fy, a5), fmaf(a7, fy, a8), fx, thr_u, thr_v, nu, nv)) return false;
  sx = (int)(nu - kRoundMagicBits) + base_x;
  sy = (int)(nv - kRoundMagicBits) + base_y;
  if (flags != 0u && (((flags & kEdgeLeft) && sx < 32 + 2) || ((flags & kEdgeRight) && sx > 32 * (W - 2) - 2) ||
                      ((flags & kEdgeTop) && sy < 32 + 2) || ((flags & kEdgeBottom) && sy > 32 * (H - 2) - 2)))
    return false;
  return true;
}


// ---------------------------------------------------------------------------------------------
// Crop edges from row segments (mfs.py:1075-1098 without touching a pixel).
//
// The reference scans the float32 maps of a frame for the largest column with |map_x| < 1, the smallest
// column with |map_x - (W-1)| < 1, the largest row with |map_y| < 1 and the smallest row with
// |map_y - (H-1)| < 1.  A pixel's map comes from the cell that owns it, and on one output row a cell owns
// whole intervals (row segments).  Along a row the map of one cell is (a x + b) / (h6 x + d) with a
// denominator of constant sign (kMapMonotone), hence MONOTONE in x: the pixels inside an open band
// (m_lo, m_hi) form one interval, whose ends follow in closed form.  Only integers within one pixel of
// that real interval can satisfy the band: the float64 evaluation differs from real arithmetic by
// ~1e-13 and rounding to float32 snaps any value that close to a band limit ONTO the limit (the limits
// -1, 1, W-2, W, H-2, H are float32 numbers and rounding is monotone), where the strict comparison
// fails either way.  Every candidate is decided by the reference's exact float64 -> float32 sequence
// (map_row), so the result equals the scan over all pixels.  Cells without kMapMonotone are scanned.
// ---------------------------------------------------------------------------------------------
struct RowMapEval {
  double h0, h2, h3, h5, h6, yh1, yh4, yh7;
  MF_HD void init(const Cell& c, int y) {
    const double yd = (double)y;
    h0 = c.Hsu[0]; h2 = c.Hsu[2]; h3 = c.Hsu[3]; h5 = c.Hsu[5]; h6 = c.Hsu[6];
    yh1 = MF_MUL(yd, c.Hsu[1]); yh4 = MF_MUL(yd, c.Hsu[4]); yh7 = MF_MUL(yd, c.Hsu[7]);
  }
  MF_HD float at(int x, bool use_y) const {
    float mx, my;
    map_row(h0, h2, h3, h5, h6, (double)x, yh1, yh4, yh7, mx, my);
    return use_y ? my : mx;
  }
};

// mode 0: largest x in [xa, xb] whose coordinate lies in (m_lo, m_hi); 1: smallest such x; 2: any such x.
// Returns -1 when there is none.
MF_HD int band_search(const RowMapEval& ev, bool use_y, float m_lo, float m_hi, int xa, int xb, int mode, bool monotone) {
  if (xa > xb) return -1;
  int A = xa, B = xb;
  if (monotone) {
    const float va = ev.at(xa, use_y), vb = ev.at(xb, use_y);
    const bool ina = va > m_lo && va < m_hi, inb = vb > m_lo && vb < m_hi;
    if (mode == 0) { if (inb) return xb; }
    else { if (ina) return xa; if (mode == 2 && inb) return xb; }
    if ((va <= m_lo && vb <= m_lo) || (va >= m_hi && vb >= m_hi)) return -1;      // monotone: nothing in between either
    const double a = use_y ? ev.h3 : ev.h0;
    const double b = use_y ? (ev.yh4 + ev.h5) : (ev.yh1 + ev.h2);
    const double d = ev.yh7 + 1.0;
    const double x1 = ((double)m_lo * d - b) / (a - (double)m_lo * ev.h6);
    const double x2 = ((double)m_hi * d - b) / (a - (double)m_hi * ev.h6);
    if (fabs(x1) < 1e8 && fabs(x2) < 1e8) {                                    // false for inf / NaN
      const double lo = x1 < x2 ? x1 : x2, hi = x1 < x2 ? x2 : x1;
      const int ca = (int)floor(lo) - 1, cb = (int)ceil(hi) + 1;
      A = ca > xa ? ca : xa;
      B = cb < xb ? cb : xb;
    }
  }
  if (mode == 0) {
    for (int x = B; x >= A; --x) { const float v = ev.at(x, use_y); if (v > m_lo && v < m_hi) return x; }
  } else {
    for (int x = A; x <= B; ++x) { const float v = ev.at(x, use_y); if (v > m_lo && v < m_hi) return x; }
  }
  return -1;
}

// Row segment [xa, xb] of row y owned by cell c: fold its pixels into the four edge searches.
// e[0] = left (max), e[1] = top (max), e[2] = right (min), e[3] = bottom (min): running values, also used
// to skip work that cannot improve them.
MF_HD void segment_crop_edges(const Cell& c, int xa, int xb, int y, int W, int H, int* e) {
  const unsigned fl = c.edge_flags;
  if (!(fl & kEdgeAny) || xa > xb) return;
  const bool mono = (fl & kMapMonotone) != 0u;
  RowMapEval ev;
  ev.init(c, y);
  if (fl & kEdgeLeft) {
    const int x = band_search(ev, false, -1.0f, 1.0f, xa > e[0] + 1 ? xa : e[0] + 1, xb, 0, mono);
    if (x > e[0]) e[0] = x;
  }
  if (fl & kEdgeRight) {
    const int x = band_search(ev, false, (float)(W - 2), (float)W, xa, xb < e[2] - 1 ? xb : e[2] - 1, 1, mono);
    if (x >= 0 && x < e[2]) e[2] = x;
  }
  if ((fl & kEdgeTop) && y > e[1]) {
    if (band_search(ev, true, -1.0f, 1.0f, xa, xb, 2, mono) >= 0) e[1] = y;
  }
  if ((fl & kEdgeBottom) && y < e[3]) {
    if (band_search(ev, true, (float)(H - 2), (float)H, xa, xb, 2, mono) >= 0) e[3] = y;
  }
}

}  // namespace mf
