// dem_contact.cuh -- device-side contact model chain (surface -> normal -> tangential -> rolling)
// for sphere/sphere and sphere/wall contacts, written for sm_100a.
//
// Behavioural contract (what must match the reference; file:line relative to the reference src/):
//   surface  default : surface_model_default.h:146-211
//   normal   hertz   : normal_model_hertz.h:205-266,366-383
//   normal   hooke   : normal_model_hooke.h:230-300
//   tangential history: tangential_model_history.h:136-240,288-334,404-440
//   rolling  cdt     : rolling_model_cdt.h:91-167
//   rolling  epsd    : rolling_model_epsd.h:97-340 ; epsd2: rolling_model_epsd2.h:152-205
//   chain order      : contact_models.h:228-238 (history slots: tangential, then rolling)
// Everything lives in registers; the caller owns loads/stores of the history row.
#pragma once
#include "dem_types.h"

namespace dem {

struct Contact {
  // geometry, oriented "first - second"
  double dx, dy, dz, r, rinv, radi, radj, radsum, deltan_in;
  double meff, mi, mj;
  double vi[3], vj[3], wi[3], wj[3];
  int itype, jtype;
};

struct ContactOut {
  double F[3];   // force on the first body (second gets the exact negative)
  double Ti[3];  // torque on the first body
  double Tj[3];  // torque on the second body
};

__device__ __forceinline__ double tabv(const StepP &P, int which, int it, int jt)
{
  return __ldg(P.tab + (which * P.nt1 + it) * P.nt1 + jt);
}
// ONE = single atom type: the material constants are kernel parameters (uniform), no per-thread loads
template <bool ONE>
__device__ __forceinline__ double tabp(const StepP &P, int which, int tij)
{
  return ONE ? P.t1[which] : __ldg(P.tab + which * P.nt1 * P.nt1 + tij);
}

// shear / ch: the tangential shear vector and the rolling spring torque of this pair's history row
// (fixed-size so that they live in registers; the caller maps them to rows off_shear.. / off_roll..)
template <int NORMAL, int ROLLING, bool WALL>
__device__ __forceinline__ void contact_chain(const StepP &P, const ModelP &M, const Contact &c,
                                              double (&shear)[3], double (&ch)[3], bool shearupdate, ContactOut &o)
{
  const double enx = c.dx * c.rinv, eny = c.dy * c.rinv, enz = c.dz * c.rinv;
  // ---- surface model: relative kinematics at the contact point
  const double vr1 = c.vi[0] - c.vj[0], vr2 = c.vi[1] - c.vj[1], vr3 = c.vi[2] - c.vj[2];
  const double vn = vr1 * enx + vr2 * eny + vr3 * enz;
  const double vt1 = vr1 - vn * enx, vt2 = vr2 - vn * eny, vt3 = vr3 - vn * enz;
  const double deltan = c.radsum - c.r;
  double wr1, wr2, wr3, cri, crj = 0.0;
  if (WALL) {
    cri = c.radi - 0.5 * c.deltan_in;  // the wall driver's overlap, not radsum - r
    wr1 = cri * c.wi[0] * c.rinv; wr2 = cri * c.wi[1] * c.rinv; wr3 = cri * c.wi[2] * c.rinv;
  } else {
    cri = c.radi - 0.5 * deltan; crj = c.radj - 0.5 * deltan;
    wr1 = (cri * c.wi[0] + crj * c.wj[0]) * c.rinv;
    wr2 = (cri * c.wi[1] + crj * c.wj[1]) * c.rinv;
    wr3 = (cri * c.wi[2] + crj * c.wj[2]) * c.rinv;
  }
  const double vtr1 = vt1 - (c.dz * wr2 - c.dy * wr3);
  const double vtr2 = vt2 - (c.dx * wr3 - c.dz * wr1);
  const double vtr3 = vt3 - (c.dy * wr1 - c.dx * wr2);

  // ---- normal model
  const double reff = WALL ? c.radi : (c.radi * c.radj / (c.radi + c.radj));
  const double meff = c.meff;
  double kn, kt, gamman, gammat;
  if (NORMAL == N_HERTZ) {
    const double Y = tabv(P, T_YEFF, c.itype, c.jtype), G = tabv(P, T_GEFF, c.itype, c.jtype);
    const double beta = tabv(P, T_BETA, c.itype, c.jtype);
    const double sqrtval = sqrt(reff * deltan);
    const double Sn = 2. * Y * sqrtval, St = 8. * G * sqrtval;
    kn = 4. / 3. * Y * sqrtval; kt = St;
    const double sqrtFiveOverSix = 0.91287092917527685576161630466800355658790782499663875;
    gamman = -2. * sqrtFiveOverSix * beta * sqrt(Sn * meff);
    gammat = M.tdamp ? -2. * sqrtFiveOverSix * beta * sqrt(St * meff) : 0.0;
  } else {
    const double Y = tabv(P, T_YEFF, c.itype, c.jtype);
    const double lg = tabv(P, T_CORLOG, c.itype, c.jtype);
    const double sqrtval = sqrt(reff);
    kn = 16. / 15. * sqrtval * Y * pow(15. * meff * P.charVel * P.charVel / (16. * sqrtval * Y), 0.2);
    kt = kn;
    if (M.ktToKn) kt *= 0.285714286;
    const double lgsq = lg * lg;
    gamman = sqrt(4. * meff * kn * lgsq / (lgsq + 3.14159265358979323846 * 3.14159265358979323846));
    gammat = M.tdamp ? gamman : 0.0;
  }
  kn /= P.nktv2p; kt /= P.nktv2p;
  double Fn = -gamman * vn + kn * deltan;
  if (M.limitForce && Fn < 0.0) Fn = 0.0;
  o.F[0] = Fn * enx; o.F[1] = Fn * eny; o.F[2] = Fn * enz;
  o.Ti[0] = o.Ti[1] = o.Ti[2] = 0.0; o.Tj[0] = o.Tj[1] = o.Tj[2] = 0.0;

  // ---- tangential model: history
  if (M.tangential) {
    if (shearupdate) {
      shear[0] += vtr1 * P.dt; shear[1] += vtr2 * P.dt; shear[2] += vtr3 * P.dt;
      const double rsht = shear[0] * enx + shear[1] * eny + shear[2] * enz;
      shear[0] -= rsht * enx; shear[1] -= rsht * eny; shear[2] -= rsht * enz;
    }
    const double shrmag = sqrt(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2]);
    const double xmu = tabv(P, T_MU, c.itype, c.jtype);
    double Ft1 = -(kt * shear[0]), Ft2 = -(kt * shear[1]), Ft3 = -(kt * shear[2]);
    const double Ft_shear = kt * shrmag, Ft_friction = xmu * fabs(Fn);
    if (Ft_shear > Ft_friction) {
      if (shrmag != 0.0) {
        const double ratio = Ft_friction / Ft_shear;
        Ft1 *= ratio; Ft2 *= ratio; Ft3 *= ratio;
        if (shearupdate) { shear[0] = -Ft1 / kt; shear[1] = -Ft2 / kt; shear[2] = -Ft3 / kt; }
      } else Ft1 = Ft2 = Ft3 = 0.0;
    } else {
      Ft1 -= gammat * vtr1; Ft2 -= gammat * vtr2; Ft3 -= gammat * vtr3;
    }
    const double tor1 = eny * Ft3 - enz * Ft2, tor2 = enz * Ft1 - enx * Ft3, tor3 = enx * Ft2 - eny * Ft1;
    o.F[0] += Ft1; o.F[1] += Ft2; o.F[2] += Ft3;
    o.Ti[0] += -cri * tor1; o.Ti[1] += -cri * tor2; o.Ti[2] += -cri * tor3;
    if (!WALL) { o.Tj[0] += -crj * tor1; o.Tj[1] += -crj * tor2; o.Tj[2] += -crj * tor3; }
  }

  // ---- rolling friction
  if (ROLLING == R_CDT) {
    const double rmu = tabv(P, T_RMU, c.itype, c.jtype);
    double a1, a2, a3;
    if (WALL) { a1 = wr1; a2 = wr2; a3 = wr3; }
    else { a1 = c.wi[0] - c.wj[0]; a2 = c.wi[1] - c.wj[1]; a3 = c.wi[2] - c.wj[2]; }
    const double mag = sqrt(a1 * a1 + a2 * a2 + a3 * a3);
    if (mag > 0.) {
      double r1, r2, r3;
      if (WALL) {
        const double FnS = deltan * kn;
        r1 = rmu * FnS * a1 / mag * reff; r2 = rmu * FnS * a2 / mag * reff; r3 = rmu * FnS * a3 / mag * reff;
      } else {
        const double sc = rmu * kn * deltan * reff / mag;
        r1 = a1 * sc; r2 = a2 * sc; r3 = a3 * sc;
      }
      if (!M.torsion) {
        const double dot = r1 * enx + r2 * eny + r3 * enz;
        r1 -= enx * dot; r2 -= eny * dot; r3 -= enz * dot;
      }
      o.Ti[0] -= r1; o.Ti[1] -= r2; o.Ti[2] -= r3;
      o.Tj[0] += r1; o.Tj[1] += r2; o.Tj[2] += r3;
    }
  } else if (ROLLING == R_EPSD || ROLLING == R_EPSD2) {
    double a1, a2, a3, r_inertia;
    if (WALL) {
      a1 = wr1; a2 = wr2; a3 = wr3;
      r_inertia = 1.4 * c.mi * reff * reff;
    } else {
      a1 = c.wi[0] - c.wj[0]; a2 = c.wi[1] - c.wj[1]; a3 = c.wi[2] - c.wj[2];
      const double ri = c.mi * c.radi * c.radi, rj = c.mj * c.radj * c.radj;
      r_inertia = 1.4 * ri * rj / (ri + rj);
    }
    const double rmu = tabv(P, T_RMU, c.itype, c.jtype);
    double w1 = a1, w2 = a2, w3 = a3;
    if (!M.torsion) {
      const double dot = a1 * enx + a2 * eny + a3 * enz;
      w1 = a1 - enx * dot; w2 = a2 - eny * dot; w3 = a3 - enz * dot;
    }
    const double kr = (ROLLING == R_EPSD2) ? kt * reff * reff : 2.25 * kn * rmu * rmu * reff * reff;
    double r1 = ch[0] + w1 * (P.dt * kr), r2 = ch[1] + w2 * (P.dt * kr), r3 = ch[2] + w3 * (P.dt * kr);
    const double mag = sqrt(r1 * r1 + r2 * r2 + r3 * r3);
    const double tmax = fabs(Fn) * reff * rmu;
    if (mag > tmax) {
      const double factor = tmax / mag;
      r1 *= factor; r2 *= factor; r3 *= factor;
      if (shearupdate) { ch[0] = r1; ch[1] = r2; ch[2] = r3; }
    } else {
      if (shearupdate) { ch[0] = r1; ch[1] = r2; ch[2] = r3; }
      if (ROLLING == R_EPSD) {
        const double r_coef = tabv(P, T_RVISC, c.itype, c.jtype) * 2 * sqrt(r_inertia * kr);
        r1 += r_coef * w1; r2 += r_coef * w2; r3 += r_coef * w3;
      }
    }
    o.Ti[0] -= r1; o.Ti[1] -= r2; o.Ti[2] -= r3;
    o.Tj[0] += r1; o.Tj[1] += r2; o.Tj[2] += r3;
  }
}


// ---------------------------------------------------------------------------------------------
// fast fp64 reciprocal / square root: MUFU seed (rcp.approx / rsqrt.approx, ~2^-20) + two Newton
// rounds, <= 1 ulp.  The correctly rounded CUDA sequences cost ~3x the instructions and the
// contact math below is issue/latency bound, not HBM bound (profiles/r01*).  Errors stay ~1e-16,
// six orders below the 1e-10 force tolerance of the parity tests.
__device__ __forceinline__ double rcp_fast(double x)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0); y = fma(y, e, y);
  e = fma(-x, y, 1.0); y = fma(y, e, y);
  return y;
}
// s = sqrt(x), rs = 1/sqrt(x); x == 0 gives s = rs = 0
__device__ __forceinline__ void sqrt_rsqrt_fast(double x, double &s, double &rs)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx, y * y, 0.5); y = fma(y, e, y);
  e = fma(-hx, y * y, 0.5); y = fma(y, e, y);
  double t = x * y;
  t = fma(fma(-t, t, x), 0.5 * y, t);
  const bool ok = x > 0.0;
  s = ok ? t : 0.0; rs = ok ? y : 0.0;
}

// Sphere/sphere contact evaluated by EACH of the two owners in its own orientation (me - partner).
// Every expression below is exactly antisymmetric (vectors) or symmetric (scalars) under the
// exchange of the two bodies -- the only sum of two cross-body products, wr, is written with
// explicit un-contracted roundings -- so the two evaluations of a pair are exact mirror images:
// forces are equal and opposite to the last bit and the two history copies stay exact negatives,
// without selecting operands into a canonical order.  shear/ch are in MY orientation.
// Adds the force / torque acting on me to F / T.
template <int NORMAL, int ROLLING, bool ONE>
__device__ __forceinline__ void pair_chain(const StepP &P, const ModelP &M, const double4 &xi, const double4 &vi, const double4 &wi,
                                           const double4 &xj, const double4 &vj, const double4 &wj, int itype, int jtype,
                                           int imask, int jmask, double dx, double dy, double dz, double rsq,
                                           double (&shear)[3], double (&ch)[3], bool shearupdate, double *F, double *T)
{
  const int tij = itype * P.nt1 + jtype;
  double r, rinv;
  sqrt_rsqrt_fast(rsq, r, rinv);
  const double enx = dx * rinv, eny = dy * rinv, enz = dz * rinv;
  const double radi = xi.w, radj = xj.w, mi = vi.w, mj = vj.w;
  const double radsum = radi + radj;
  // surface model (surface_model_default.h:146-211)
  const double vr1 = vi.x - vj.x, vr2 = vi.y - vj.y, vr3 = vi.z - vj.z;
  const double vn = vr1 * enx + vr2 * eny + vr3 * enz;
  const double vt1 = vr1 - vn * enx, vt2 = vr2 - vn * eny, vt3 = vr3 - vn * enz;
  const double deltan = radsum - r;
  const double cri = radi - 0.5 * deltan, crj = radj - 0.5 * deltan;
  const double wr1 = __dadd_rn(__dmul_rn(cri, wi.x), __dmul_rn(crj, wj.x)) * rinv;
  const double wr2 = __dadd_rn(__dmul_rn(cri, wi.y), __dmul_rn(crj, wj.y)) * rinv;
  const double wr3 = __dadd_rn(__dmul_rn(cri, wi.z), __dmul_rn(crj, wj.z)) * rinv;
  const double vtr1 = vt1 - (dz * wr2 - dy * wr3);
  const double vtr2 = vt2 - (dx * wr3 - dz * wr1);
  const double vtr3 = vt3 - (dy * wr1 - dx * wr2);
  // effective radius / mass (pair_gran_base.h:386-393)
  const double reff = radi * radj * rcp_fast(radsum);
  double meff = mi * mj * rcp_fast(mi + mj);
  if (imask & P.freezebit) meff = mj;
  if (jmask & P.freezebit) meff = mi;
  double kn, kt, inv_kt, gamman, gammat;
  if (NORMAL == N_HERTZ) {  // normal_model_hertz.h:205-266
    const double Y = tabp<ONE>(P, T_YEFF, tij), G = tabp<ONE>(P, T_GEFF, tij);
    const double beta = tabp<ONE>(P, T_BETA, tij);
    double s, inv_s, q, inv_q;
    sqrt_rsqrt_fast(reff * deltan, s, inv_s);
    sqrt_rsqrt_fast(s * meff, q, inv_q);
    kn = 4. / 3. * Y * s; kt = 8. * G * s;
    inv_kt = tabp<ONE>(P, T_INV8G, tij) * inv_s;
    const double c2 = -2. * 0.91287092917527685576161630466800355658790782499663875 * beta * q;
    gamman = c2 * tabp<ONE>(P, T_SQ2Y, tij);          // -2 sqrt(5/6) beta sqrt(Sn meff), Sn = 2 Y s
    gammat = M.tdamp ? c2 * tabp<ONE>(P, T_SQ8G, tij) : 0.0;  // St = 8 G s
  } else {  // normal_model_hooke.h:230-300
    const double Y = tabp<ONE>(P, T_YEFF, tij);
    const double lg = tabp<ONE>(P, T_CORLOG, tij);
    const double sqrtval = sqrt(reff);
    kn = 16. / 15. * sqrtval * Y * pow(15. * meff * P.charVel * P.charVel / (16. * sqrtval * Y), 0.2);
    kt = kn;
    if (M.ktToKn) kt *= 0.285714286;
    const double lgsq = lg * lg;
    gamman = sqrt(4. * meff * kn * lgsq / (lgsq + 3.14159265358979323846 * 3.14159265358979323846));
    gammat = M.tdamp ? gamman : 0.0;
    inv_kt = P.nktv2p / kt;
  }
  if (P.nktv2p != 1.0) { kn /= P.nktv2p; kt /= P.nktv2p; if (NORMAL == N_HERTZ) inv_kt *= P.nktv2p; }
  double Fn = -gamman * vn + kn * deltan;
  if (M.limitForce && Fn < 0.0) Fn = 0.0;
  double F1 = Fn * enx, F2 = Fn * eny, F3 = Fn * enz;
  double T1 = 0.0, T2 = 0.0, T3 = 0.0;
  if (M.tangential) {  // tangential_model_history.h:136-240,288-334
    if (shearupdate) {
      shear[0] += vtr1 * P.dt; shear[1] += vtr2 * P.dt; shear[2] += vtr3 * P.dt;
      const double rsht = shear[0] * enx + shear[1] * eny + shear[2] * enz;
      shear[0] -= rsht * enx; shear[1] -= rsht * eny; shear[2] -= rsht * enz;
    }
    double shrmag, inv_shr;
    sqrt_rsqrt_fast(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2], shrmag, inv_shr);
    const double xmu = tabp<ONE>(P, T_MU, tij);
    double Ft1 = -(kt * shear[0]), Ft2 = -(kt * shear[1]), Ft3 = -(kt * shear[2]);
    const double Ft_shear = kt * shrmag, Ft_friction = xmu * fabs(Fn);
    if (Ft_shear > Ft_friction) {
      if (shrmag != 0.0) {
        const double ratio = Ft_friction * (inv_kt * inv_shr);
        Ft1 *= ratio; Ft2 *= ratio; Ft3 *= ratio;
        if (shearupdate) { shear[0] = -Ft1 * inv_kt; shear[1] = -Ft2 * inv_kt; shear[2] = -Ft3 * inv_kt; }
      } else Ft1 = Ft2 = Ft3 = 0.0;
    } else {
      Ft1 -= gammat * vtr1; Ft2 -= gammat * vtr2; Ft3 -= gammat * vtr3;
    }
    F1 += Ft1; F2 += Ft2; F3 += Ft3;
    T1 = -cri * (eny * Ft3 - enz * Ft2); T2 = -cri * (enz * Ft1 - enx * Ft3); T3 = -cri * (enx * Ft2 - eny * Ft1);
  }
  if (ROLLING != R_OFF) {
    const double a1 = wi.x - wj.x, a2 = wi.y - wj.y, a3 = wi.z - wj.z;
    const double rmu = tabp<ONE>(P, T_RMU, tij);
    if (ROLLING == R_CDT) {  // rolling_model_cdt.h:91-167
      double mag, inv_mag;
      sqrt_rsqrt_fast(a1 * a1 + a2 * a2 + a3 * a3, mag, inv_mag);
      if (mag > 0.) {
        const double sc = rmu * kn * deltan * reff * inv_mag;
        double r1 = a1 * sc, r2 = a2 * sc, r3 = a3 * sc;
        if (!M.torsion) {
          const double dot = r1 * enx + r2 * eny + r3 * enz;
          r1 -= enx * dot; r2 -= eny * dot; r3 -= enz * dot;
        }
        T1 -= r1; T2 -= r2; T3 -= r3;
      }
    } else {  // rolling_model_epsd.h:97-340, rolling_model_epsd2.h:152-205
      double w1 = a1, w2 = a2, w3 = a3;
      if (!M.torsion) {
        const double dot = a1 * enx + a2 * eny + a3 * enz;
        w1 = a1 - enx * dot; w2 = a2 - eny * dot; w3 = a3 - enz * dot;
      }
      const double kr = (ROLLING == R_EPSD2) ? kt * reff * reff : 2.25 * kn * rmu * rmu * reff * reff;
      double r1 = ch[0] + w1 * (P.dt * kr), r2 = ch[1] + w2 * (P.dt * kr), r3 = ch[2] + w3 * (P.dt * kr);
      double mag, inv_mag;
      sqrt_rsqrt_fast(r1 * r1 + r2 * r2 + r3 * r3, mag, inv_mag);
      const double tmax = fabs(Fn) * reff * rmu;
      if (mag > tmax) {
        const double factor = tmax * inv_mag;
        r1 *= factor; r2 *= factor; r3 *= factor;
        if (shearupdate) { ch[0] = r1; ch[1] = r2; ch[2] = r3; }
      } else {
        if (shearupdate) { ch[0] = r1; ch[1] = r2; ch[2] = r3; }
        if (ROLLING == R_EPSD) {
          const double ri = mi * radi * radi, rj = mj * radj * radj;
          const double r_inertia = 1.4 * ri * rj * rcp_fast(ri + rj);
          const double r_coef = tabp<ONE>(P, T_RVISC, tij) * 2 * sqrt(r_inertia * kr);
          r1 += r_coef * w1; r2 += r_coef * w2; r3 += r_coef * w3;
        }
      }
      T1 -= r1; T2 -= r2; T3 -= r3;
    }
  }
  F[0] += F1; F[1] += F2; F[2] += F3;
  T[0] += T1; T[1] += T2; T[2] += T3;
}

}  // namespace dem
