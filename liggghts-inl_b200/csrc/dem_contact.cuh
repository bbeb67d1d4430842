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

// hist: dnum doubles (tangential shear at off_shear, rolling spring torque at off_roll)
template <int NORMAL, int ROLLING, bool WALL>
__device__ __forceinline__ void contact_chain(const StepP &P, const ModelP &M, const Contact &c,
                                              double *hist, bool shearupdate, ContactOut &o)
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
    double *shear = hist + M.off_shear;
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
    double *ch = hist + M.off_roll;
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

}  // namespace dem
