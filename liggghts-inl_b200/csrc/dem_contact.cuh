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
  double *nh;  // history of the normal model (hysteretic/nonlinear1|2: 12 values), else unused
  double *th;  // tangential model hysteretic/nonlinear: [0] = shrmag_0 (history value 3; values 4..6 are read, never written: 0)
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

// INL normal laws hysteretic/nonlinear1 and 2 (normal_model_hysteretic_nonlinear1.h:149-385, normal_model_hysteretic_nonlinear2.h):
// piecewise loading / unloading / reloading force with plastic overlap, 12 history values per pair (deltaMax, deltaZero, k1,
// deltaZero_old, k1_old, delta_old, deltaMin, betan, f0, f0_old, kc, fo) that the model rewrites in every force evaluation,
// the setup one included.  Returns Fn; kn, kt (= k1), gamman, gammat go to the tangential / rolling models.
template <bool V2>
__device__ __forceinline__ double hyst_normal(const StepP &P, const ModelP &M, const Contact &c, double deltan, double vn,
                                              double &kn_o, double &kt_o, double &gamman_o, double &gammat_o)
{
  const int it = c.itype, jt = c.jtype;
  const double PI = 3.14159265358979323846;
  const double meff = c.meff;
  const double Alpha = tabv(P, T_H_ALPHA, it, jt), Cin = tabv(P, T_H_CIN, it, jt), A1 = tabv(P, T_H_A1, it, jt), A2 = tabv(P, T_H_A2, it, jt);
  const double A3 = tabv(P, T_H_A3, it, jt);
  const double k_c = tabv(P, T_H_KCIN, it, jt) * A2;
  const double kc = tabv(P, T_H_KN2KC, it, jt) * tabv(P, T_H_KEL, it, jt);
  const double f_0 = tabv(P, T_H_FADH, it, jt);
  const double crl = tabv(P, T_CORLOG, it, jt);
  const double cdamp = 1. + (PI / crl) * (PI / crl);
  double *h = c.nh;
  double deltaMax;
  if (deltan > h[0]) { h[0] = deltan; deltaMax = deltan; } else deltaMax = h[0];
  double deltaZero = h[1], k1 = h[2], deltaZero_old = h[3];
  const double k1_old = h[4], delta_old = h[5];
  double deltaMin = h[6], betan = h[7], f0 = h[8], f0_old = h[9];
  double k2 = A3 * k1, fHys, gamman;
  const bool loading = deltan >= delta_old ? (delta_old == 0 ? true : !(vn > 0)) : false;
  const double dexp = V2 ? 0.5 : 0.25;
  if (loading) {
    if (deltaZero == 0) k1 = A2;
    f0_old = f0;
    if (deltan <= deltaZero) {
      if (!V2) fHys = deltan <= deltaMin ? -k_c * deltan : betan * (deltan - deltaZero);
      else fHys = Cin * k2 * (exp(betan * (deltan - deltaZero)) - 1);
    } else fHys = Alpha * k1 * pow(deltan - deltaZero, 2.0) + f0;
    gamman = sqrt(4. * meff * Alpha * k1 / cdamp) * (pow(deltan, dexp) + pow(deltaZero, dexp));  // (sqrt(5/4) is sqrt(1): integer division)
    h[1] = deltaZero; h[2] = k1; h[3] = deltaZero; h[4] = k1; h[5] = deltan; h[6] = deltaMin; h[7] = betan; h[8] = f0; h[9] = f0_old;
  } else if (!V2) {
    k2 = A3 * k1_old;
    deltaZero = (1 - k1_old / k2) * deltaMax;
    const double beta = Alpha * k1_old * pow(deltaMax - deltaZero_old, 2.0) / k2 / (deltaMax - deltaZero);
    deltaMin = beta * (k2 - k1_old) / (beta * k2 + k_c) * deltaMax;
    k1 = deltaMax * A1 + A2;
    if (deltan >= deltaMin) {
      betan = beta * k2;
      if (deltan >= deltaZero) { fHys = beta * k2 * (deltan - deltaZero) + (deltan - deltaZero) * f0_old / (deltaMax - deltaZero); f0 = fHys - Alpha * k1 * pow(deltan - deltaZero, 2.0); }
      else { fHys = beta * k2 * (deltan - deltaZero); f0 = 0; }
    } else { fHys = -k_c * deltan; f0 = 0; }
    gamman = sqrt(4. * meff * Alpha * k1_old / cdamp) * pow(deltan, 0.25);
    h[1] = deltaZero; h[2] = k1; h[3] = deltaZero_old; h[4] = k1_old; h[5] = deltan; h[6] = deltaMin; h[7] = betan; h[8] = f0;
  } else {
    k2 = A3 * (A1 * deltaMax + A2);
    deltaZero = (1 - k1_old / k2) * deltaMax;
    betan = log(Alpha * k1_old / Cin / k2 * pow(deltaMax - deltaZero_old, 2.0) + 1) / (deltaMax - (1 - k1_old / k2) * deltaMax);
    deltaMin = betan * (k2 - k1_old) / (betan * k2 + k_c) * deltaMax;
    k1 = deltaMax * A1 + A2;
    if (deltan >= deltaZero) { fHys = Cin * k2 * (exp(betan * (deltan - deltaZero)) - 1) + (deltan - deltaZero) * f0_old / (deltaMax - deltaZero); f0 = fHys - Alpha * k1 * pow(deltan - deltaZero, 2.0); }
    else { fHys = Cin * k2 * (exp(betan * (deltan - deltaZero)) - 1); f0 = 0; }
    gamman = 1 * 0.001 * sqrt(4. * meff * Alpha * k1_old / cdamp) * pow(deltan, -0.25);
    h[1] = deltaZero; h[2] = k1; h[3] = deltaZero_old; h[4] = k1_old; h[5] = deltan; h[6] = deltaMin; h[7] = betan; h[8] = f0;
  }
  double Fn = fHys + (-gamman * vn) + f_0;
  if (M.limitForce && Fn < 0.0 && kc == 0 && f_0 == 0.0) Fn = 0.0;
  h[10] = kc; h[11] = f_0;
  kn_o = k1 / P.nktv2p; kt_o = k1 / P.nktv2p; gamman_o = gamman; gammat_o = gamman;  // (tangential_damping is registered but never applied)
  return Fn;
}

// shear / ch: the tangential shear vector and the rolling spring torque of this pair's history row
// (fixed-size so that they live in registers; the caller maps them to rows off_shear.. / off_roll..)
template <int NORMAL, int ROLLING, bool WALL>
__device__ __forceinline__ void contact_chain(const StepP &P, const ModelP &M, const Contact &c,
                                              double (&shear)[3], double (&ch)[3], bool shearupdate, ContactOut &o, bool drop_normal = false)
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
  } else if (NORMAL == N_HOOKE) {
    const double Y = tabv(P, T_YEFF, c.itype, c.jtype);
    const double lg = tabv(P, T_CORLOG, c.itype, c.jtype);
    const double sqrtval = sqrt(reff);
    kn = 16. / 15. * sqrtval * Y * pow(15. * meff * P.charVel * P.charVel / (16. * sqrtval * Y), 0.2);
    kt = kn;
    if (M.ktToKn) kt *= 0.285714286;
    const double lgsq = lg * lg;
    gamman = sqrt(4. * meff * kn * lgsq / (lgsq + 3.14159265358979323846 * 3.14159265358979323846));
    gammat = M.tdamp ? gamman : 0.0;
  } else { kn = kt = gamman = gammat = 0.0; }
  double Fn;
  if (NORMAL == N_HYST1 || NORMAL == N_HYST2) Fn = hyst_normal<NORMAL == N_HYST2>(P, M, c, deltan, vn, kn, kt, gamman, gammat);
  else {
    kn /= P.nktv2p; kt /= P.nktv2p;
    Fn = -gamman * vn + kn * deltan;
    if (M.limitForce && Fn < 0.0) Fn = 0.0;
  }
  o.F[0] = Fn * enx; o.F[1] = Fn * eny; o.F[2] = Fn * enz;
  if (drop_normal) o.F[0] = o.F[1] = o.F[2] = 0.0;  // cohesion bond/nonlinear assigns the pair force after the normal model
  o.Ti[0] = o.Ti[1] = o.Ti[2] = 0.0; o.Tj[0] = o.Tj[1] = o.Tj[2] = 0.0;

  // ---- tangential model: history
  if (M.tangential) {
    if (shearupdate) {
      shear[0] += vtr1 * P.dt; shear[1] += vtr2 * P.dt; shear[2] += vtr3 * P.dt;
      const double rsht = shear[0] * enx + shear[1] * eny + shear[2] * enz;
      shear[0] -= rsht * enx; shear[1] -= rsht * eny; shear[2] -= rsht * enz;
    }
    double shrmag = sqrt(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2]);
    const double xmu = tabv(P, T_MU, c.itype, c.jtype);
    if ((NORMAL == N_HYST1 || NORMAL == N_HYST2) && M.tangential == 2) {
      // tangential model hysteretic/nonlinear (tangential_model_hysteretic_nonlinear.h:186-206): while the overlap is inside the
      // plastic range of the normal law (deltan <= deltaZero) the spring restarts and the magnitude is remembered
      double shrmag_0 = c.th[0];
      if (deltan <= c.nh[1]) { shrmag_0 = shrmag; c.th[0] = shrmag_0; shear[0] = 0.0; shear[1] = 0.0; shear[2] = 0.0; }
      shrmag -= shrmag_0;
    }
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
        const double FnS = M.cdtnl2 ? Fn : deltan * kn;  // cdtnonlinear2: rolling_model_cdtnonlinear2.h:127
        r1 = rmu * FnS * a1 / mag * reff; r2 = rmu * FnS * a2 / mag * reff; r3 = rmu * FnS * a3 / mag * reff;
      } else {
        const double sc = M.cdtnl2 ? rmu * Fn * reff / mag : rmu * kn * deltan * reff / mag;  // cdtnonlinear2: :157
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
template <int NORMAL, int ROLLING, bool ONE, bool STD = false>
__device__ __forceinline__ void pair_chain(const StepP &P, const ModelP &M, const double4 &xi, const double4 &vi, const double4 &wi,
                                           const double4 &xj, const double4 &vj, const double4 &wj, int itype, int jtype,
                                           int imask, int jmask, double dx, double dy, double dz, double rsq,
                                           double (&shear)[3], double (&ch)[3], bool shearupdate, double *F, double *T, double *Tp = nullptr)
{  // Tp (optional, owner list): receives the torque on the PARTNER, -crj (en x Ft) + rolling torque
  // STD: the reference's default sub-model settings, known at compile time (see pair_item in dem_kernels.cuh)
  const bool m_tangential = STD ? true : (M.tangential != 0), m_tdamp = STD ? true : (M.tdamp != 0);
  const bool m_limitForce = STD ? false : (M.limitForce != 0), m_torsion = STD ? false : (M.torsion != 0);
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
    gammat = m_tdamp ? c2 * tabp<ONE>(P, T_SQ8G, tij) : 0.0;  // St = 8 G s
  } else {  // normal_model_hooke.h:230-300
    const double Y = tabp<ONE>(P, T_YEFF, tij);
    const double lg = tabp<ONE>(P, T_CORLOG, tij);
    const double sqrtval = sqrt(reff);
    kn = 16. / 15. * sqrtval * Y * pow(15. * meff * P.charVel * P.charVel / (16. * sqrtval * Y), 0.2);
    kt = kn;
    if (M.ktToKn) kt *= 0.285714286;
    const double lgsq = lg * lg;
    gamman = sqrt(4. * meff * kn * lgsq / (lgsq + 3.14159265358979323846 * 3.14159265358979323846));
    gammat = m_tdamp ? gamman : 0.0;
    inv_kt = P.nktv2p / kt;
  }
  if (!STD && P.nktv2p != 1.0) { kn /= P.nktv2p; kt /= P.nktv2p; if (NORMAL == N_HERTZ) inv_kt *= P.nktv2p; }
  double Fn = -gamman * vn + kn * deltan;
  if (m_limitForce && Fn < 0.0) Fn = 0.0;
  double F1 = Fn * enx, F2 = Fn * eny, F3 = Fn * enz;
  double T1 = 0.0, T2 = 0.0, T3 = 0.0, P1 = 0.0, P2 = 0.0, P3 = 0.0;
  if (m_tangential) {  // tangential_model_history.h:136-240,288-334
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
    if (Tp) { P1 = -crj * (eny * Ft3 - enz * Ft2); P2 = -crj * (enz * Ft1 - enx * Ft3); P3 = -crj * (enx * Ft2 - eny * Ft1); }
  }
  if (ROLLING != R_OFF) {
    const double a1 = wi.x - wj.x, a2 = wi.y - wj.y, a3 = wi.z - wj.z;
    const double rmu = tabp<ONE>(P, T_RMU, tij);
    if (ROLLING == R_CDT) {  // rolling_model_cdt.h:91-167
      double mag, inv_mag;
      sqrt_rsqrt_fast(a1 * a1 + a2 * a2 + a3 * a3, mag, inv_mag);
      if (mag > 0.) {
        const double sc = (!STD && M.cdtnl2) ? rmu * Fn * reff * inv_mag : rmu * kn * deltan * reff * inv_mag;  // cdtnonlinear2: rolling_model_cdtnonlinear2.h:157
        double r1 = a1 * sc, r2 = a2 * sc, r3 = a3 * sc;
        if (!m_torsion) {
          const double dot = r1 * enx + r2 * eny + r3 * enz;
          r1 -= enx * dot; r2 -= eny * dot; r3 -= enz * dot;
        }
        T1 -= r1; T2 -= r2; T3 -= r3;
        P1 += r1; P2 += r2; P3 += r3;
      }
    } else {  // rolling_model_epsd.h:97-340, rolling_model_epsd2.h:152-205
      double w1 = a1, w2 = a2, w3 = a3;
      if (!m_torsion) {
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
      P1 += r1; P2 += r2; P3 += r3;
    }
  }
  F[0] += F1; F[1] += F2; F[2] += F3;
  T[0] += T1; T[1] += T2; T[2] += T3;
  if (Tp) { Tp[0] += P1; Tp[1] += P2; Tp[2] += P3; }
}


// ---------------------------------------------------------------------------------------------
// fp32 mode (option `fp32`; BASELINE.json north_star: forces within 1e-5 of the fp64 reference): the same chain with the
// contact LAW in single precision.  What stays in fp64: the particle state, the separation r, the unit normal and the
// overlap radsum - r (the subtraction cancels seven to eight digits -- an overlap formed in fp32 would carry a relative
// error of ~3e-5 and the Hertz force 1.5 times that), relative velocities (formed in fp64, then rounded), the force / torque
// sums and the integration.  History values are advanced in fp32 and stored as doubles.  Like pair_chain the expressions are
// exactly (anti)symmetric under exchange of the bodies, so the two owners of a pair still produce mirror images.
__device__ __forceinline__ float rsqrt_f32(float x)
{
  const float y = rsqrtf(x);
  return x > 0.f ? y * (1.5f - 0.5f * x * y * y) : 0.f;
}
template <int NORMAL, int ROLLING, bool ONE>
__device__ __forceinline__ void pair_chain_f32(const StepP &P, const ModelP &M, const double4 &xi, const double4 &vi, const double4 &wi,
                                               const double4 &xj, const double4 &vj, const double4 &wj, int itype, int jtype,
                                               int imask, int jmask, double dx, double dy, double dz, double rsq,
                                               double (&shear)[3], double (&ch)[3], bool shearupdate, double *F, double *T, double *Tp)
{
  const int tij = itype * P.nt1 + jtype;
  double r, rinv;
  sqrt_rsqrt_fast(rsq, r, rinv);
  const float enx = (float)(dx * rinv), eny = (float)(dy * rinv), enz = (float)(dz * rinv);
  const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz, frinv = (float)rinv;
  const float radi = (float)xi.w, radj = (float)xj.w, mi = (float)vi.w, mj = (float)vj.w;
  const float deltan = (float)((xi.w + xj.w) - r);
  const float dt = (float)P.dt;
  const float vr1 = (float)(vi.x - vj.x), vr2 = (float)(vi.y - vj.y), vr3 = (float)(vi.z - vj.z);
  const float vn = vr1 * enx + vr2 * eny + vr3 * enz;
  const float vt1 = vr1 - vn * enx, vt2 = vr2 - vn * eny, vt3 = vr3 - vn * enz;
  const float cri = radi - 0.5f * deltan, crj = radj - 0.5f * deltan;
  const float wix = (float)wi.x, wiy = (float)wi.y, wiz = (float)wi.z, wjx = (float)wj.x, wjy = (float)wj.y, wjz = (float)wj.z;
  const float wr1 = __fadd_rn(__fmul_rn(cri, wix), __fmul_rn(crj, wjx)) * frinv;
  const float wr2 = __fadd_rn(__fmul_rn(cri, wiy), __fmul_rn(crj, wjy)) * frinv;
  const float wr3 = __fadd_rn(__fmul_rn(cri, wiz), __fmul_rn(crj, wjz)) * frinv;
  const float vtr1 = vt1 - (fdz * wr2 - fdy * wr3);
  const float vtr2 = vt2 - (fdx * wr3 - fdz * wr1);
  const float vtr3 = vt3 - (fdy * wr1 - fdx * wr2);
  const float reff = radi * radj / (radi + radj);
  float meff = mi * mj / (mi + mj);
  if (imask & P.freezebit) meff = mj;
  if (jmask & P.freezebit) meff = mi;
  const float nk = (float)P.nktv2p;
  float kn, kt, inv_kt, gamman, gammat;
  if (NORMAL == N_HERTZ) {
    const float Y = (float)tabp<ONE>(P, T_YEFF, tij), G = (float)tabp<ONE>(P, T_GEFF, tij), beta = (float)tabp<ONE>(P, T_BETA, tij);
    const float rd = reff * deltan;
    const float inv_s = rsqrt_f32(rd), s = rd * inv_s;
    const float q = sqrtf(s * meff);
    kn = 4.f / 3.f * Y * s; kt = 8.f * G * s;
    inv_kt = (float)tabp<ONE>(P, T_INV8G, tij) * inv_s;
    const float c2 = -2.f * 0.912870929175276855761616f * beta * q;
    gamman = c2 * (float)tabp<ONE>(P, T_SQ2Y, tij);
    gammat = M.tdamp ? c2 * (float)tabp<ONE>(P, T_SQ8G, tij) : 0.f;
  } else {
    const float Y = (float)tabp<ONE>(P, T_YEFF, tij), lg = (float)tabp<ONE>(P, T_CORLOG, tij);
    const float sqrtval = sqrtf(reff), cv = (float)P.charVel;
    kn = 16.f / 15.f * sqrtval * Y * powf(15.f * meff * cv * cv / (16.f * sqrtval * Y), 0.2f);
    kt = kn;
    if (M.ktToKn) kt *= 0.285714286f;
    const float lgsq = lg * lg;
    gamman = sqrtf(4.f * meff * kn * lgsq / (lgsq + 9.86960440108935861883f));
    gammat = M.tdamp ? gamman : 0.f;
    inv_kt = nk / kt;
  }
  if (P.nktv2p != 1.0) { kn /= nk; kt /= nk; if (NORMAL == N_HERTZ) inv_kt *= nk; }
  float Fn = -gamman * vn + kn * deltan;
  if (M.limitForce && Fn < 0.f) Fn = 0.f;
  float F1 = Fn * enx, F2 = Fn * eny, F3 = Fn * enz;
  float T1 = 0.f, T2 = 0.f, T3 = 0.f, P1 = 0.f, P2 = 0.f, P3 = 0.f;
  if (M.tangential) {
    float s0 = (float)shear[0], s1 = (float)shear[1], s2 = (float)shear[2];
    if (shearupdate) {
      s0 += vtr1 * dt; s1 += vtr2 * dt; s2 += vtr3 * dt;
      const float rsht = s0 * enx + s1 * eny + s2 * enz;
      s0 -= rsht * enx; s1 -= rsht * eny; s2 -= rsht * enz;
    }
    const float ssq = s0 * s0 + s1 * s1 + s2 * s2;
    const float inv_shr = rsqrt_f32(ssq), shrmag = ssq * inv_shr;
    const float xmu = (float)tabp<ONE>(P, T_MU, tij);
    float Ft1 = -(kt * s0), Ft2 = -(kt * s1), Ft3 = -(kt * s2);
    const float Ft_shear = kt * shrmag, Ft_friction = xmu * fabsf(Fn);
    if (Ft_shear > Ft_friction) {
      if (shrmag != 0.f) {
        const float ratio = Ft_friction * (inv_kt * inv_shr);
        Ft1 *= ratio; Ft2 *= ratio; Ft3 *= ratio;
        if (shearupdate) { s0 = -Ft1 * inv_kt; s1 = -Ft2 * inv_kt; s2 = -Ft3 * inv_kt; }
      } else Ft1 = Ft2 = Ft3 = 0.f;
    } else {
      Ft1 -= gammat * vtr1; Ft2 -= gammat * vtr2; Ft3 -= gammat * vtr3;
    }
    if (shearupdate) { shear[0] = s0; shear[1] = s1; shear[2] = s2; }
    F1 += Ft1; F2 += Ft2; F3 += Ft3;
    const float c1 = eny * Ft3 - enz * Ft2, c2 = enz * Ft1 - enx * Ft3, c3 = enx * Ft2 - eny * Ft1;
    T1 = -cri * c1; T2 = -cri * c2; T3 = -cri * c3;
    if (Tp) { P1 = -crj * c1; P2 = -crj * c2; P3 = -crj * c3; }
  }
  if (ROLLING != R_OFF) {
    const float a1 = (float)(wi.x - wj.x), a2 = (float)(wi.y - wj.y), a3 = (float)(wi.z - wj.z);
    const float rmu = (float)tabp<ONE>(P, T_RMU, tij);
    if (ROLLING == R_CDT) {
      const float asq = a1 * a1 + a2 * a2 + a3 * a3;
      if (asq > 0.f) {
        const float sc = (M.cdtnl2 ? rmu * Fn * reff : rmu * kn * deltan * reff) * rsqrt_f32(asq);
        float r1 = a1 * sc, r2 = a2 * sc, r3 = a3 * sc;
        if (!M.torsion) {
          const float dot = r1 * enx + r2 * eny + r3 * enz;
          r1 -= enx * dot; r2 -= eny * dot; r3 -= enz * dot;
        }
        T1 -= r1; T2 -= r2; T3 -= r3;
        P1 += r1; P2 += r2; P3 += r3;
      }
    } else {
      float w1 = a1, w2 = a2, w3 = a3;
      if (!M.torsion) {
        const float dot = a1 * enx + a2 * eny + a3 * enz;
        w1 = a1 - enx * dot; w2 = a2 - eny * dot; w3 = a3 - enz * dot;
      }
      const float kr = (ROLLING == R_EPSD2) ? kt * reff * reff : 2.25f * kn * rmu * rmu * reff * reff;
      float r1 = (float)ch[0] + w1 * (dt * kr), r2 = (float)ch[1] + w2 * (dt * kr), r3 = (float)ch[2] + w3 * (dt * kr);
      const float msq = r1 * r1 + r2 * r2 + r3 * r3;
      const float inv_mag = rsqrt_f32(msq), mag = msq * inv_mag;
      const float tmax = fabsf(Fn) * reff * rmu;
      if (mag > tmax) {
        const float factor = tmax * inv_mag;
        r1 *= factor; r2 *= factor; r3 *= factor;
        if (shearupdate) { ch[0] = r1; ch[1] = r2; ch[2] = r3; }
      } else {
        if (shearupdate) { ch[0] = r1; ch[1] = r2; ch[2] = r3; }
        if (ROLLING == R_EPSD) {
          const float ri = mi * radi * radi, rj = mj * radj * radj;
          const float r_inertia = 1.4f * ri * rj / (ri + rj);
          const float r_coef = (float)tabp<ONE>(P, T_RVISC, tij) * 2.f * sqrtf(r_inertia * kr);
          r1 += r_coef * w1; r2 += r_coef * w2; r3 += r_coef * w3;
        }
      }
      T1 -= r1; T2 -= r2; T3 -= r3;
      P1 += r1; P2 += r2; P3 += r3;
    }
  }
  F[0] += (double)F1; F[1] += (double)F2; F[2] += (double)F3;
  T[0] += (double)T1; T[1] += (double)T2; T[2] += (double)T3;
  if (Tp) { Tp[0] += (double)P1; Tp[1] += (double)P2; Tp[2] += (double)P3; }
}

// ---------------------------------------------------------------------------------------------
// Bonded-sphere models, sphere-sphere branch:
//   cohesion bond            cohesion_model_bond.h:491-950 (createBond :1015-1034, breakBond :1036-1070)
//   cohesion bond/nonlinear  cohesion_model_bond_nonlinear.h:394-940 (:958-995)
// Evaluated in the canonical orientation "first body = lower tag" by both owners of a pair, with identical operands in
// identical order, so that the two copies of the history stay bit-identical (the running max/min trackers of the nonlinear
// model are orientation dependent and carry no newtonflag in the reference: a canonical orientation is the only consistent
// choice).  H = the nbond history doubles of the pair (bondFlag, initial_dist, contactPos[3], ft[3], tn[3], tt[3], trackers).
// Returns true when the bond acts (the reference's has_force_update); F/Ti/Tj = force on the first body, torques on both.
__device__ __forceinline__ void vproject(const double *v, const double *on, double *res)
{  // vector_liggghts.h:437-442: `on` is normalised first (zero vector -> zero)
  const double norm = sqrt(on[0] * on[0] + on[1] * on[1] + on[2] * on[2]);
  const double inv = (norm == 0.) ? 0. : 1. / norm;
  const double n0 = on[0] * inv, n1 = on[1] * inv, n2 = on[2] * inv;
  const double d = v[0] * n0 + v[1] * n1 + v[2] * n2;
  res[0] = n0 * d; res[1] = n1 * d; res[2] = n2 * d;
}
__device__ __forceinline__ double dsgn(double v) { return (double)((0. < v) - (v < 0.)); }
__device__ __forceinline__ double damp_mult(const ModelP &M, double vel, double minvel, double dv)
{
  return M.dampingSmooth ? fmin(1.0, fmax(-1.0, vel / fmax(minvel, dv))) : dsgn(vel);
}
__device__ __forceinline__ double nl_torque_comp(double theta, double &tmax, double &tmin, double k, double ku, double kc, double JI)
{  // one component of the hysteretic twist / bend law, cohesion_model_bond_nonlinear.h:669-872
  if (theta > 0.0) tmin = 0.0; else if (theta < 0.0) tmax = 0.0;
  const double c1 = (ku - k) / (ku + kc) * tmax, c2 = (ku - k) / (ku + kc) * tmin;
  double tq;
  if ((theta >= tmax) || (theta <= tmin)) tq = -k * JI * theta;
  else if (theta > c1) tq = -ku * JI * theta + (ku - k) * JI * tmax;
  else if (theta >= c2) tq = kc * JI * theta;
  else tq = -ku * JI * theta + (ku - k) * JI * tmin;
  if (theta > tmax) tmax = theta;
  if (theta < tmin) tmin = theta;
  return tq;
}

template <int COH>
__device__ __forceinline__ bool bond_eval(const StepP &P, const ModelP &M, const double *delta, double rsq, double radi, double radj,
                                       const double *xi, const double *vi, const double *vj, const double *omegai, const double *omegaj,
                                       int it, int jt, bool update_history, double *H, double *F, double *Ti, double *Tj, int &events)
{  // events: bit 0 a bond was created, bit 1 a bond broke (compute bond/counter, cohesion_model_bond.h:1032,1066)
  constexpr bool NL = (COH == C_BONDNL);
  const double lambda = tabv(P, T_B_LAMBDA, it, jt);
  if (lambda < 1.e-15) return false;
  const double r = sqrt(rsq);
  if (update_history) {
    if (H[0] < 1.e-15) {
      const bool create = (M.createAlways || (P.ntimestep == P.tsCreateBond)) && (r < tabv(P, T_B_CREATEDIST, it, jt));
      if (!create) return false;
      H[0] = 1.0; H[1] = r;
      for (int d = 0; d < 3; d++) H[2 + d] = xi[d] - delta[d];
      for (int d = 5; d < 14; d++) H[d] = 0.0;
      events |= 1;
    }
  } else if (H[0] < 1.e-15) return false;
  double force_tang[3] = {H[5], H[6], H[7]}, tn[3] = {H[8], H[9], H[10]}, tt[3] = {H[11], H[12], H[13]};
  if (!M.stressBreak && r > tabv(P, T_B_MAXDIST, it, jt) && update_history) { H[0] = 0.; H[1] = 0.; events |= 2; return false; }
  const double rinv = 1. / r;
  const double en[3] = {delta[0] * rinv, delta[1] * rinv, delta[2] * rinv};
  const double rb = lambda * (radi < radj ? radi : radj);
  const double A = 3.14159265358979323846 * rb * rb, J = 0.5 * A * rb * rb, I = 0.5 * J, dt = P.dt;
  const double radsuminv = 1. / (radi + radj);
  const double cri = r * radi * radsuminv, crj = r * radj * radsuminv;
  double vr[3], vn[3], vt[3], wr[3], tmp1[3], vtr[3], wn[3], wt[3];
  for (int d = 0; d < 3; d++) vr[d] = vi[d] - vj[d];
  vproject(vr, en, vn);
  for (int d = 0; d < 3; d++) vt[d] = vr[d] - vn[d];
  for (int d = 0; d < 3; d++) wr[d] = omegai[d] * (radi * radsuminv) + omegaj[d] * (radj * radsuminv);
  tmp1[0] = delta[1] * wr[2] - delta[2] * wr[1]; tmp1[1] = delta[2] * wr[0] - delta[0] * wr[2]; tmp1[2] = delta[0] * wr[1] - delta[1] * wr[0];
  for (int d = 0; d < 3; d++) vtr[d] = vt[d] + tmp1[d];
  for (int d = 0; d < 3; d++) wr[d] = omegai[d] - omegaj[d];
  vproject(wr, en, wn);
  for (int d = 0; d < 3; d++) wt[d] = wr[d] - wn[d];
  double nforce[3] = {0., 0., 0.}, nforce_d[3] = {0., 0., 0.}, tforce_d[3] = {0., 0., 0.}, ntorque_d[3] = {0., 0., 0.}, ttorque_d[3] = {0., 0., 0.};
  double torque_normal[3] = {0., 0., 0.}, torque_tang[3] = {0., 0., 0.};
  const double displacement = H[1] - r;
  const double minvel = 1e-5 * fmin(radi, radj) / dt;
  const double dfn = tabv(P, T_B_DFN, it, jt), dft = tabv(P, T_B_DFT, it, jt), dtn = tabv(P, T_B_DTN, it, jt), dtt = tabv(P, T_B_DTT, it, jt);
  double dmax = 0., dmin = 0.;
  if (NL) {  // running extreme displacements, updated even when shearupdate == 0 (:566-580)
    if (displacement > H[14]) { H[14] = displacement; dmax = displacement; } else dmax = H[14];
    if (displacement < H[27]) { H[27] = displacement; dmin = displacement; } else dmin = H[27];
    if (displacement < 0.0) H[14] = 0.0;
    if (displacement > 0.0) H[27] = 0.0;
  }
  if (M.tension || M.compression) {
    if (!NL && M.dissipation && update_history) H[1] += (r - H[1]) * fmin(dt * tabv(P, T_D_FN, it, jt), 1.0);  // relax the normal spring (:677-687; the force below uses the displacement formed before)
    if ((M.tension && displacement < -1.e-15) || (M.compression && displacement > 1.e-15)) {
      double frcmag;
      if (!NL) frcmag = tabv(P, T_B_KN, it, jt) * A * displacement;
      else {
        const double k1 = tabv(P, T_B_K_FN1, it, jt), ku1 = tabv(P, T_B_KU_FN1, it, jt), kc1 = tabv(P, T_B_KC_FN1, it, jt);
        const double k2 = tabv(P, T_B_K_FN2, it, jt), ku2 = tabv(P, T_B_KU_FN2, it, jt), kc2 = tabv(P, T_B_KC_FN2, it, jt);
        const double q1 = (ku1 - k1) / (ku1 + kc1);
        const double c1 = (q1 * q1) * dmax, c2 = ((ku2 - k2) / (ku2 + kc2)) * dmin;  // pow(x,2.0) and pow(x,1.0) are exact products in libm
        if (displacement < dmin) frcmag = k2 * A * displacement;
        else if (displacement < c2) frcmag = ku2 * A * displacement + (k2 - ku2) * A * dmin;
        else if (displacement < 0.0) frcmag = -kc2 * A * displacement;
        else if (displacement < c1) frcmag = -kc1 * sqrt(fabs(displacement));
        else if (displacement < dmax) frcmag = ku1 * sqrt(fabs(displacement)) + (k1 - ku1) * sqrt(fabs(dmax));
        else frcmag = k1 * sqrt(fabs(displacement));
      }
      for (int d = 0; d < 3; d++) nforce[d] = en[d] * frcmag;
      if (M.damping) for (int d = 0; d < 3; d++) nforce_d[d] = nforce[d] - dfn * fabs(nforce[d]) * damp_mult(M, vn[d], minvel, 0.01 * nforce[d] * dt);
      else for (int d = 0; d < 3; d++) nforce_d[d] = nforce[d];
    }
  }
  if (M.shearf) {
    const double ktA = NL ? tabv(P, T_B_K_FT, it, jt) : tabv(P, T_B_KT, it, jt);
    double dtforce[3];
    for (int d = 0; d < 3; d++) dtforce[d] = vtr[d] * (-ktA * A * dt);
    vproject(force_tang, en, tmp1);
    for (int d = 0; d < 3; d++) force_tang[d] = force_tang[d] - tmp1[d];
    if (!NL && M.dissipation) { const double k = 1.0 - fmin(dt * tabv(P, T_D_FT, it, jt), 1.0); for (int d = 0; d < 3; d++) force_tang[d] = force_tang[d] * k; }  // :731-732
    for (int d = 0; d < 3; d++) force_tang[d] = force_tang[d] + dtforce[d];
    if (M.damping) for (int d = 0; d < 3; d++) tforce_d[d] = force_tang[d] - dft * fabs(force_tang[d]) * damp_mult(M, vtr[d], minvel, 0.01 * force_tang[d] * dt);
    else for (int d = 0; d < 3; d++) tforce_d[d] = force_tang[d];
  }
  if (!NL) {
    const double kn_pb = tabv(P, T_B_KN, it, jt), kt_pb = tabv(P, T_B_KT, it, jt);
    for (int d = 0; d < 3; d++) { torque_normal[d] = tn[d]; torque_tang[d] = tt[d]; }
    if (M.ntorque) {
      double dnt[3];
      for (int d = 0; d < 3; d++) dnt[d] = wn[d] * (-kt_pb * J * dt);
      vproject(torque_normal, en, torque_normal);
      if (M.dissipation) { const double k = 1.0 - fmin(dt * tabv(P, T_D_TN, it, jt), 1.0); for (int d = 0; d < 3; d++) torque_normal[d] = torque_normal[d] * k; }  // :763-764
      for (int d = 0; d < 3; d++) torque_normal[d] = torque_normal[d] + dnt[d];
      if (M.damping) for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d] - dtn * fabs(torque_normal[d]) * dsgn(wn[d]);
      else for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d];
    }
    if (M.ttorque) {
      const double wtsq = wt[0] * wt[0] + wt[1] * wt[1] + wt[2] * wt[2];
      if (wtsq > 0) {
        double dtt3[3];
        for (int d = 0; d < 3; d++) dtt3[d] = wt[d] * (-kn_pb * I * dt);
        vproject(torque_tang, wt, torque_tang);
        if (M.dissipation) { const double k = 1.0 - fmin(dt * tabv(P, T_D_TT, it, jt), 1.0); for (int d = 0; d < 3; d++) torque_tang[d] = torque_tang[d] * k; }  // :798-799
        for (int d = 0; d < 3; d++) torque_tang[d] = torque_tang[d] + dtt3[d];
        if (M.damping) for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d] - dtt * fabs(torque_tang[d]) * dsgn(wt[d]);
        else for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d];
      }
    }
  } else {
    if (M.ntorque) {
      const double k = tabv(P, T_B_K_TN, it, jt), ku = tabv(P, T_B_KU_TN, it, jt), kc = tabv(P, T_B_KC_TN, it, jt);
      for (int d = 0; d < 3; d++) tn[d] = tn[d] + wn[d] * dt;
      for (int d = 0; d < 3; d++) torque_normal[d] = nl_torque_comp(tn[d], H[15 + d], H[21 + d], k, ku, kc, J);
      if (M.damping) for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d] - dtn * fabs(torque_normal[d]) * dsgn(wn[d]);
      else for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d];
    }
    if (M.ttorque) {
      const double k = tabv(P, T_B_K_TT, it, jt), ku = tabv(P, T_B_KU_TT, it, jt), kc = tabv(P, T_B_KC_TT, it, jt);
      for (int d = 0; d < 3; d++) tt[d] = tt[d] + wt[d] * dt;
      for (int d = 0; d < 3; d++) torque_tang[d] = nl_torque_comp(tt[d], H[18 + d], H[24 + d], k, ku, kc, I);
      if (M.damping) for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d] - dtt * fabs(torque_tang[d]) * dsgn(wt[d]);
      else for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d];
    }
  }
  if (M.stressBreak) {  // un-damped forces / torques (:816-848, :875-890)
    const double nfm = sqrt(nforce[0] * nforce[0] + nforce[1] * nforce[1] + nforce[2] * nforce[2]);
    const double tfm = sqrt(force_tang[0] * force_tang[0] + force_tang[1] * force_tang[1] + force_tang[2] * force_tang[2]);
    const double ntm = sqrt(torque_normal[0] * torque_normal[0] + torque_normal[1] * torque_normal[1] + torque_normal[2] * torque_normal[2]);
    const double ttm = sqrt(torque_tang[0] * torque_tang[0] + torque_tang[1] * torque_tang[1] + torque_tang[2] * torque_tang[2]);
    double maxSigma = tabv(P, T_B_MAXSIGMA, it, jt);
    if (M.ratioTC && (NL ? displacement < -1.e-15 : displacement < 1e-16)) maxSigma *= tabv(P, T_B_RATIOTC, it, jt);
    const bool nstress = maxSigma < (nfm / A + ttm * rb / I);
    const bool tstress = tabv(P, T_B_MAXTAU, it, jt) < (tfm / A + ntm * rb / J);
    if ((nstress || tstress) && update_history) { H[0] = 0.; H[1] = 0.; events |= 2; return false; }
  }
  const double tor[3] = {tforce_d[1] * en[2] - tforce_d[2] * en[1], tforce_d[2] * en[0] - tforce_d[0] * en[2], tforce_d[0] * en[1] - tforce_d[1] * en[0]};
  for (int d = 0; d < 3; d++) { F[d] = nforce_d[d] + tforce_d[d]; Ti[d] = cri * tor[d] + ntorque_d[d] + ttorque_d[d]; Tj[d] = crj * tor[d] - ntorque_d[d] - ttorque_d[d]; }
  if (update_history) for (int d = 0; d < 3; d++) { H[5 + d] = force_tang[d]; H[8 + d] = NL ? tn[d] : torque_normal[d]; H[11 + d] = NL ? tt[d] : torque_tang[d]; }
  return true;
}

}  // namespace dem
