// weno_fast.cuh — the FAST-math reconstruction of the burgers stage (device; also compiles
// as plain C++ so tests/test_fast_math_cpu.py can check it against the oracle on the CPU).
//
// WENO5-Z + MC-limited linear blend of benchmarks/burgers/recon.hpp:27-99, reorganised to
// issue ~66 FP64 instructions (+ ~8 for the differences a marching thread shares between
// neighbouring cells) instead of the ~165 of the reference's expression tree:
//   * everything is written in the four first differences d_k = q_k - q_{k-1}: the
//     smoothness indicators need one op per term instead of two, the candidate values
//     become offsets from q2 (two ops each, and q2 is added once at the very end inside
//     the final FMA chain), and the limiter reuses d2, d3 and d2 + d3;
//   * the curvature part of a smoothness indicator is shared by three neighbouring cells
//     (weno_curv: computed once per row by a marching thread);
//   * the 8 divisions collapse into 2 reciprocals: 1/(b0 b1 b2) for the three ratios r_k, and
//     one common denominator for both sides' normalisations and the blend weight (see
//     WENO5Z_diff); each is a hardware seed (MUFU.RCP64H, 20+ bits) plus ONE cubically
//     convergent step (3 FMAs, error e^3 < 2^-60);
//   * the limiter's sign test (dm dp > 0) is an integer test on the sign bits.
// Results stay within ~1e-14 of the reference expression relative to the stencil's magnitude
// (<= 1e-12 relative on evolved fields is the contract: tests/test_burgers_sim_gpu.py pins it at
// the benchmark's cycle count).  Valid while products of three smoothness indicators do not
// overflow (|q| < ~1e40).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PB2_HD __device__ __forceinline__
#else
#define PB2_HD inline
#endif

namespace pb2 {
namespace fastmath {

PB2_HD double min_std(double a, double b) { return (b < a) ? b : a; }
PB2_HD double max_std(double a, double b) { return (a < b) ? b : a; }

// 1/a: hardware seed + one third-order step  x = x0 (1 + e + e^2),  e = 1 - a x0
PB2_HD double rcp_fast(double a) {
#if defined(__CUDA_ARCH__)
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);
  const double t = fma(e, e, e);
  return fma(x, t, x);
#else
  // host stand-in for the hardware seed: the exact quotient cut to 20 mantissa bits
  double x = 1.0 / a;
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  bits &= ~((uint64_t(1) << 32) - 1);
  std::memcpy(&x, &bits, 8);
  const double e = std::fma(-a, x, 1.0);
  const double t = std::fma(e, e, e);
  return std::fma(x, t, x);
#endif
}

PB2_HD bool same_sign(double a, double b) {
#if defined(__CUDA_ARCH__)
  return (__double2hiint(a) ^ __double2hiint(b)) >= 0;
#else
  int64_t x, y;
  std::memcpy(&x, &a, 8);
  std::memcpy(&y, &b, 8);
  return (x ^ y) >= 0;
#endif
}

PB2_HD int64_t dbits(double a) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(a);
#else
  int64_t x;
  std::memcpy(&x, &a, 8);
  return x;
#endif
}
PB2_HD double from_dbits(int64_t x) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(x);
#else
  double a;
  std::memcpy(&a, &x, 8);
  return a;
#endif
}

// 0.5 * mc(dm, dp) of recon.hpp:27-32 given s = dm + dp:
//   copysign(min(|s|/4, min(|dm|, |dp|)), s) when dm and dp have the same sign, else 0
// (exact: only powers of two are moved through the min; a zero difference gives 0 either way).
// PB2_INT_LIMITER = 1 does it entirely on the bit patterns (magnitudes of doubles order like
// integers, |s|/4 is an exponent decrement; |s| < 2^-1020 is flushed to 0) so that the limiter
// costs no FP64-pipe instruction.
// A/B on one B200 (profiles/README.md, session r02l): the integer limiter LOSES 1.2 % (the extra
// integer instructions cost more issue slots than the three FP64 instructions they replace
// free on the pipe), the power-of-two scaling is neutral to +0.5 %: off / on by default.
#ifndef PB2_INT_LIMITER
#define PB2_INT_LIMITER 0
#endif
#ifndef PB2_POW2_SCALE
#define PB2_POW2_SCALE 1
#endif
PB2_HD double half_mc(double dm, double dp, double s) {
#if !PB2_INT_LIMITER
  const double mm = min_std(min_std(fabs(dm), fabs(dp)), 0.25 * fabs(s));
  return same_sign(dm, dp) ? copysign(mm, s) : 0.0;
#else
  constexpr int64_t kMag = 0x7fffffffffffffffll;
  const int64_t bm = dbits(dm), bp = dbits(dp), bs = dbits(s);
  const int64_t am = bm & kMag, ap = bp & kMag;
  int64_t as = bs & kMag;
  as = as >= (int64_t(3) << 52) ? as - (int64_t(2) << 52) : 0;
  int64_t m = am < ap ? am : ap;
  m = as < m ? as : m;
  return from_dbits((bm ^ bp) >= 0 ? (m | (bs & ~kMag)) : 0);
#endif
}

PB2_HD void Linear(const double qm, const double q0, const double qp, double &ql, double &qr) {
  const double dm = q0 - qm, dp = qp - q0;
  const double dq = half_mc(dm, dp, dm + dp);
  ql = q0 + dq;
  qr = q0 - dq;
}

// Constants that do not fit the short-immediate form of an FP64 instruction live in constant
// memory, so they are read as c[bank][offset] operands instead of being rebuilt with two moves
// in front of every use.
struct WenoConst {
  double c13;           // 13/3
  double g0, g1, g2;    // linear weights 0.1, 0.6, 0.3
  double h0, h1, h2;    // g / 6
  double eps3;          // eps / 3
};
#if defined(__CUDACC__)
__constant__ WenoConst kW = {13.0 / 3.0, 0.1, 0.6, 0.3, 0.1 / 6.0, 0.6 / 6.0, 0.3 / 6.0,
                             10.0 * DBL_EPSILON / 3.0};
#else
static const WenoConst kW = {13.0 / 3.0, 0.1, 0.6, 0.3, 0.1 / 6.0, 0.6 / 6.0, 0.3 / 6.0,
                             10.0 * DBL_EPSILON / 3.0};
#endif

// c13 a^2 + eps of the second difference a = dhi - dlo: the part of a smoothness indicator that
// THREE neighbouring cells share (beta0 of cell c+1, beta1 of cell c, beta2 of cell c-1 all hold
// the second difference centred at c) — a marching thread computes it once per row
PB2_HD double weno_curv(const double dlo, const double dhi) {
  constexpr double eps = 10.0 * DBL_EPSILON; // robust.hpp:39-42
  const double a = dhi - dlo;
  return fma(kW.c13 * a, a, eps);
}

// WENO5-Z of cell c from its four first differences d_k = q_k - q_{k-1} (d1 = q1 - q0 ... d4 =
// q4 - q3 in the numbering of recon.hpp:43), the curvature terms A0, A1, A2 = weno_curv centred
// at c-1, c, c+1 and the cell value q2
PB2_HD void WENO5Z_diff(const double d1, const double d2, const double d3, const double d4,
                        const double A0, const double A1, const double A2, const double q2,
                        double &ql, double &qr) {
  const double g0 = kW.g0, g1 = kW.g1, g2 = kW.g2;
  const double s23 = d2 + d3; // q3 - q1

  // smoothness indicators (recon.hpp:52-60): curvature term + one-sided slope squared
  double b = fma(3.0, d2, -d1);
  const double b0 = fma(b, b, A0);
  const double b1 = fma(s23, s23, A1);
  b = fma(-3.0, d3, d4);
  const double b2 = fma(b, b, A2);
  const double tau5 = fabs(b2 - b0);

  // r_k = (b_k + tau5) / b_k = 1 + tau5 * (product of the other two) / (b0 b1 b2).  Everything
  // below is homogeneous of degree 0 in (r0, r1, r2), so the division by B = b0 b1 b2 is
  // replaced by a scaling with the power of two that brings B into [1, 2) (two multiplies and
  // three integer operations instead of a reciprocal): r_k here is B' times the reference's ratio
  const double b01 = b0 * b1, b12 = b1 * b2, b02 = b0 * b2;
  const double B = b01 * b2;
#if PB2_POW2_SCALE
  const double sc = from_dbits((int64_t(0x7fe) << 52) - (dbits(B) & (int64_t(0x7ff) << 52)));
  const double t = tau5 * sc, Bs = B * sc;
#else
  const double t = tau5 * rcp_fast(B), Bs = 1.0;
#endif
  const double r0 = fma(t, b12, Bs), r1 = fma(t, b02, Bs), r2 = fma(t, b01, Bs);

  // SIX times (candidate value - q2): the rows of w5alpha applied to the differences have
  // small integer coefficients (short immediates, and a coefficient 1 costs no multiply)
  const double e0 = fma(5.0, d2, -2.0 * d1), e1 = fma(2.0, d3, d2), e2 = fma(4.0, d3, -d4);
  const double f0 = fma(-5.0, d3, 2.0 * d4), f1 = fma(-2.0, d2, -d3), f2 = fma(-4.0, d2, d1);

  // Weights w_k = g_k r_k: the reference adds eps = 2.2e-15 to weights that are >= 0.1 (r_k >=
  // 1) and to alphas that are O(1) — below the contract's 1e-12 by two orders, left out here.
  // Without it the two sides share everything but their sums:  with left weights (g0 r0, g1 r1,
  // g2 r2) and right weights (g0 r2, g1 r1, g2 r0)
  //   g2 w0 w1 + g1 w0 w2 + g0 w1 w2 = g0 g1 g2 E,   E = r0 r1 + r0 r2 + r1 r2      (both sides)
  //   w0 w1 w2 = g0 g1 g2 r0 r1 r2                                                  (both sides)
  //   alpha_l = 3 r0 r1 r2 / (S_l E),  alpha_r = 3 r0 r1 r2 / (S_r E)       (recon.hpp:73-76,88-91)
  //   alpha_lin = 2 alpha_l alpha_r / (alpha_l + alpha_r) = 6 r0 r1 r2 / (E (S_l + S_r))
  // so ONE reciprocal, of E (S_l + S_r) S_l S_r, yields X = alpha_lin / 6 and X / S_l, X / S_r,
  // the factors of the two candidate sums:
  //   ql = q2 + (X / S_l) sum_k w_k e_k + (1 - 6 X) dq,    qr likewise with f_k and -dq
  const double w0 = g0 * r0, w1 = g1 * r1, w2 = g2 * r2, v0 = g0 * r2, v2 = g2 * r0;
  const double Sl = w0 + w1 + w2, Sr = v0 + w1 + v2;
  const double r12 = r1 * r2;
  const double E = fma(r0, r1 + r2, r12);
  const double Z = (r12 * r0) * rcp_fast((E * (Sl + Sr)) * (Sl * Sr));
  const double Xl = Z * Sr, Xr = Z * Sl; // X / S_l, X / S_r
  const double om = fma(-6.0 * Xl, Sl, 1.0); // 1 - alpha_lin
  const double dl = fma(w0, e0, fma(w1, e1, w2 * e2));
  const double dr = fma(v0, f0, fma(w1, f1, v2 * f2));
  const double dq = half_mc(d2, d3, s23);
  ql = fma(Xl, dl, fma(om, dq, q2));
  qr = fma(Xr, dr, fma(-om, dq, q2));
}

PB2_HD void WENO5Z(const double q0, const double q1, const double q2, const double q3,
                   const double q4, double &ql, double &qr) {
  const double d1 = q1 - q0, d2 = q2 - q1, d3 = q3 - q2, d4 = q4 - q3;
  WENO5Z_diff(d1, d2, d3, d4, weno_curv(d1, d2), weno_curv(d2, d3), weno_curv(d3, d4), q2, ql,
              qr);
}

// 0.5 * mc-limited linear reconstruction from the two first differences around the cell
PB2_HD void Linear_diff(const double dm, const double dp, const double q0, double &ql,
                        double &qr) {
  const double dq = half_mc(dm, dp, dm + dp);
  ql = q0 + dq;
  qr = q0 - dq;
}

} // namespace fastmath
} // namespace pb2
