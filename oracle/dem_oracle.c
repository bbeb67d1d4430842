/* oracle/dem_oracle.c -- TEST INFRASTRUCTURE.  NOT product code, never linked into or
 * called from libdem_b200.so.  Only tests/, __graft_entry__.smoke() and the cpu_baseline
 * leg of bench.py may load the library built from this file.
 *
 * A plain-C, single-thread, fp64 restatement of the LIGGGHTS-INL per-timestep particle
 * path.  Each function cites the reference file:line it follows
 * (paths relative to /root/reference/src).  Parity status: PINNED -- checked against
 * tests/golden/*.npz, which were produced by running the unmodified reference built by
 * oracle/Makefile.ref (generator: tests/golden/make_golden.py).
 *
 * Deliberate simplifications (none changes results beyond fp64 summation order):
 *  - local particle order is the upload order for the whole run (the reference re-sorts
 *    every 1000 steps, atom.cpp:1326-1420), so a pair (i<j) keeps its orientation and the
 *    mirrored per-atom partner arrays of fix_contact_history.cpp:305-425 reduce to a
 *    lookup of the old list row of i;
 *  - periodic images are taken on the fly (x_j + shift, the same single addition the
 *    reference does when it packs a ghost, atom_vec_sphere.cpp:283-293), one copy of the
 *    history per pair instead of one per owner;
 *  - neighbour search is a brute-force cell walk over bins of size >= cutneighmax.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXT 8   /* max atom types */
#define MAXW 16  /* max primitive walls */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

enum { N_HERTZ = 1, N_HOOKE = 2 };
enum { R_OFF = 0, R_CDT = 1, R_EPSD = 2, R_EPSD2 = 3 };
enum { CONTACT_NORMAL = 2, CONTACT_TANGENTIAL = 4, CONTACT_ROLLING = 16 }; /* contact_model_constants.h:64-69 (values irrelevant: only != 0 is tested) */

typedef struct {
  int normal, tangential, rolling;
  int tangential_damping, limitForce, torsionTorque, ktToKn;
  int dnum, off_shear, off_roll;
} model_t;

typedef struct {
  char id[64];
  model_t m;
  int wtype;      /* 0..2 plane x,y,z ; 3..5 cylinder x,y,z   primitive_wall_definitions.h:63-72 */
  double param[3];
  int atom_type;
  int shear, shearDim, shearAxis; double vshear, shearAxisVec[3];
  int ncand; int *cand;      /* PrimitiveWall::neighlist */
  double *hist;              /* n x dnum, fix property/atom "history_<id>" */
} wall_t;

typedef struct orc_engine {
  char err[256];
  double lo[3], hi[3], prd[3]; int periodic[3];
  int ntypes; double skin; int every, delay, check; double dt;
  double nktv2p, ftm2v;
  /* raw properties */
  double Y[MAXT + 1], nu[MAXT + 1], cor[MAXT + 1][MAXT + 1], mu[MAXT + 1][MAXT + 1],
      rmu[MAXT + 1][MAXT + 1], rvisc[MAXT + 1][MAXT + 1], charVel;
  /* derived (global_properties.cpp:428-560) */
  double Yeff[MAXT + 1][MAXT + 1], Geff[MAXT + 1][MAXT + 1], betaeff[MAXT + 1][MAXT + 1], corLog[MAXT + 1][MAXT + 1];
  model_t pm; int have_pair;
  wall_t walls[MAXW]; int nwalls;
  double g[3]; int have_gravity;
  int freezebit, integbit;
  double cdf; /* neighbor->contactDistanceFactor (1.0 without bond models) */
  /* particles */
  long n; int *tag, *type, *mask;
  double *x, *v, *f, *omega, *torque, *radius, *rmass, *density, *xhold;
  /* half list (CSR) + history */
  long *first; int *numneigh; int *jlist; signed char *jshift; int *flag; double *hist; long npairs, cap;
  long ntimestep, nbuilds; int ago; int setup_done;
} orc_engine;

static int fail(orc_engine *e, const char *msg) { snprintf(e->err, sizeof e->err, "%s", msg); return -1; }
const char *orc_last_error(const orc_engine *e) { return e->err; }

int orc_create(orc_engine **out, int device, int rank, int nranks, const void *id, void *stream)
{
  (void)device; (void)rank; (void)nranks; (void)id; (void)stream;
  orc_engine *e = (orc_engine *)calloc(1, sizeof *e);
  e->every = 1; e->delay = 0; e->check = 1; e->nktv2p = 1.0; e->ftm2v = 1.0; e->cdf = 1.0;
  e->integbit = 1; e->ntypes = 1; e->skin = 0.0;
  *out = e; return 0;
}
void orc_destroy(orc_engine *e)
{
  if (!e) return;
  free(e->tag); free(e->type); free(e->mask); free(e->x); free(e->v); free(e->f); free(e->omega);
  free(e->torque); free(e->radius); free(e->rmass); free(e->density); free(e->xhold);
  free(e->first); free(e->numneigh); free(e->jlist); free(e->jshift); free(e->flag); free(e->hist);
  for (int w = 0; w < e->nwalls; w++) { free(e->walls[w].cand); free(e->walls[w].hist); }
  free(e);
}
int orc_set_units(orc_engine *e, const char *s)
{ /* update.cpp:160-260: si, cgs, micro all have ftm2v = nktv2p = 1 */
  if (!strcmp(s, "si") || !strcmp(s, "cgs") || !strcmp(s, "micro")) { e->nktv2p = e->ftm2v = 1.0; return 0; }
  return fail(e, "units style not supported");
}
int orc_set_box(orc_engine *e, const double lo[3], const double hi[3], const int p[3])
{ for (int d = 0; d < 3; d++) { e->lo[d] = lo[d]; e->hi[d] = hi[d]; e->prd[d] = hi[d] - lo[d]; e->periodic[d] = p[d]; } return 0; }
int orc_set_ntypes(orc_engine *e, int n) { if (n < 1 || n > MAXT) return fail(e, "ntypes"); e->ntypes = n; return 0; }
int orc_set_processors(orc_engine *e, int a, int b, int c) { (void)e; return (a * b * c == 1) ? 0 : -1; }
int orc_set_neighbor(orc_engine *e, double skin, int every, int delay, int check)
{ e->skin = skin; e->every = every; e->delay = delay; e->check = check; return 0; }
int orc_set_timestep(orc_engine *e, double dt) { e->dt = dt; return 0; }

int orc_set_property(orc_engine *e, const char *name, const char *kind, const double *v, int n)
{
  const int T = e->ntypes;
  if (!strcmp(kind, "scalar")) {
    if (!strcmp(name, "characteristicVelocity")) { e->charVel = v[0]; return 0; }
    return fail(e, "unknown scalar property");
  }
  if (!strcmp(kind, "peratomtype")) {
    if (n != T) return fail(e, "peratomtype needs ntypes values");
    double *dst = !strcmp(name, "youngsModulus") ? e->Y : !strcmp(name, "poissonsRatio") ? e->nu : NULL;
    if (!dst) return fail(e, "unknown peratomtype property");
    for (int i = 0; i < T; i++) dst[i + 1] = v[i];
    return 0;
  }
  if (!strcmp(kind, "peratomtypepair")) {
    if (n != T * T) return fail(e, "peratomtypepair needs ntypes^2 values");
    double (*dst)[MAXT + 1] = !strcmp(name, "coefficientRestitution") ? e->cor
                            : !strcmp(name, "coefficientFriction") ? e->mu
                            : !strcmp(name, "coefficientRollingFriction") ? e->rmu
                            : !strcmp(name, "coefficientRollingViscousDamping") ? e->rvisc : NULL;
    if (!dst) return fail(e, "unknown peratomtypepair property");
    for (int i = 0; i < T; i++) for (int j = 0; j < T; j++) dst[i + 1][j + 1] = v[i * T + j];
    return 0;
  }
  return fail(e, "unknown property kind");
}

/* contact_models.cpp:158-260 (fixed keyword order) + Settings::registerOnOff of each model */
static int parse_model(orc_engine *e, int *pargc, const char *const **pargv, model_t *m)
{
  int argc = *pargc; const char *const *a = *pargv;
  memset(m, 0, sizeof *m); m->tangential_damping = 1;
  if (argc > 1 && !strcmp(a[0], "model")) {
    if (!strcmp(a[1], "hertz")) m->normal = N_HERTZ; else if (!strcmp(a[1], "hooke")) m->normal = N_HOOKE;
    else return fail(e, "normal model not supported");
    a += 2; argc -= 2;
  } else return fail(e, "expected 'model'");
  if (argc > 1 && !strcmp(a[0], "tangential")) {
    if (!strcmp(a[1], "history")) m->tangential = 1; else return fail(e, "tangential model not supported");
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "cohesion")) return fail(e, "cohesion model not supported by the oracle yet");
  if (argc > 1 && !strcmp(a[0], "rolling_friction")) {
    if (!strcmp(a[1], "cdt")) m->rolling = R_CDT; else if (!strcmp(a[1], "epsd")) m->rolling = R_EPSD;
    else if (!strcmp(a[1], "epsd2")) m->rolling = R_EPSD2; else if (!strcmp(a[1], "off")) m->rolling = R_OFF;
    else return fail(e, "rolling model not supported");
    a += 2; argc -= 2;
  }
  /* history slot order = model construction order: cohesion, tangential, rolling (contact_models.h:141-145) */
  m->dnum = 0; m->off_shear = m->off_roll = -1;
  if (m->tangential) { m->off_shear = m->dnum; m->dnum += 3; }
  if (m->rolling == R_EPSD || m->rolling == R_EPSD2) { m->off_roll = m->dnum; m->dnum += 3; }
  *pargc = argc; *pargv = a; return 0;
}
/* Settings::parseArguments: trailing `key on|off` pairs registered by the selected models */
static int parse_settings(orc_engine *e, int argc, const char *const *a, model_t *m)
{
  while (argc > 0) {
    int on;
    if (argc < 2) return fail(e, "unknown keyword or missing on/off");
    if (!strcmp(a[1], "on")) on = 1; else if (!strcmp(a[1], "off")) on = 0; else return fail(e, "expected on/off");
    if (!strcmp(a[0], "tangential_damping")) m->tangential_damping = on;
    else if (!strcmp(a[0], "limitForce")) m->limitForce = on;
    else if (!strcmp(a[0], "torsionTorque") && m->rolling != R_OFF) m->torsionTorque = on;
    else if (!strcmp(a[0], "ktToKnUser") && m->normal == N_HOOKE) m->ktToKn = on;
    else return fail(e, "unknown or unsupported setting");
    a += 2; argc -= 2;
  }
  return 0;
}

int orc_set_pair_style(orc_engine *e, int argc, const char *const *argv)
{
  if (parse_model(e, &argc, &argv, &e->pm)) return -1;
  if (parse_settings(e, argc, argv, &e->pm)) return -1;
  e->have_pair = 1; return 0;
}

int orc_add_wall_primitive(orc_engine *e, const char *id, int argc, const char *const *argv)
{ /* fix_wall_gran.cpp:171-330 */
  if (e->nwalls == MAXW) return fail(e, "too many walls");
  wall_t *w = &e->walls[e->nwalls]; memset(w, 0, sizeof *w);
  snprintf(w->id, sizeof w->id, "%s", id);
  if (parse_model(e, &argc, &argv, &w->m)) return -1;
  if (argc < 4 || strcmp(argv[0], "primitive") || strcmp(argv[1], "type")) return fail(e, "expected 'primitive type T <style> ...'");
  w->atom_type = atoi(argv[2]);
  static const char *names[6] = {"xplane", "yplane", "zplane", "xcylinder", "ycylinder", "zcylinder"};
  w->wtype = -1; for (int k = 0; k < 6; k++) if (!strcmp(argv[3], names[k])) w->wtype = k;
  if (w->wtype < 0) return fail(e, "unknown primitive wall style");
  int np = w->wtype < 3 ? 1 : 3; if (argc < 4 + np) return fail(e, "not enough wall args");
  for (int k = 0; k < np; k++) w->param[k] = atof(argv[4 + k]);
  argv += 4 + np; argc -= 4 + np; w->shearAxis = -1;
  while (argc > 0) {
    if (!strcmp(argv[0], "shear") && argc >= 3) {
      w->shearDim = argv[1][0] - 'x'; w->vshear = atof(argv[2]); w->shear = 1;
      int axis = w->wtype >= 3 ? w->wtype - 3 : -1;
      if (w->shearDim != axis) { w->shearAxis = axis; if (axis >= 0) w->shearAxisVec[axis] = w->vshear; }
      argv += 3; argc -= 3;
    } else break;
  }
  if (parse_settings(e, argc, argv, &w->m)) return -1;
  e->nwalls++; return 0;
}

int orc_set_gravity(orc_engine *e, double mag, const double dir[3])
{ /* fix_gravity.cpp:379-397 (style vector): acc = magnitude * dir/|dir| */
  double len = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  if (len == 0.0) return fail(e, "gravity direction vector = 0");
  for (int d = 0; d < 3; d++) { const double u = dir[d] / len; e->g[d] = mag * u; } /* fix_gravity.cpp:382-397 */
  e->have_gravity = 1; return 0;
}
int orc_set_freeze(orc_engine *e, int bit) { e->freezebit = bit; return 0; }
int orc_set_integrate(orc_engine *e, int bit) { e->integbit = bit; return 0; }

int orc_upload_particles(orc_engine *e, long n, const int *tag, const int *type, const int *mask,
                         const double *x, const double *v, const double *omega, const double *radius, const double *density)
{
  e->n = n;
#define AL(p, T, c) p = (T *)calloc((size_t)(n ? n : 1) * (c), sizeof(T))
  AL(e->tag, int, 1); AL(e->type, int, 1); AL(e->mask, int, 1); AL(e->x, double, 3); AL(e->v, double, 3);
  AL(e->f, double, 3); AL(e->omega, double, 3); AL(e->torque, double, 3); AL(e->radius, double, 1);
  AL(e->rmass, double, 1); AL(e->density, double, 1); AL(e->xhold, double, 3);
  for (long i = 0; i < n; i++) {
    e->tag[i] = tag[i]; e->type[i] = type[i]; e->mask[i] = mask ? mask[i] : 1;
    for (int d = 0; d < 3; d++) { e->x[3 * i + d] = x[3 * i + d]; e->v[3 * i + d] = v ? v[3 * i + d] : 0.0; e->omega[3 * i + d] = omega ? omega[3 * i + d] : 0.0; }
    e->radius[i] = radius[i]; e->density[i] = density[i];
    /* atom_vec_sphere.cpp:1078-1079 */
    e->rmass[i] = 4.0 * M_PI / 3.0 * radius[i] * radius[i] * radius[i] * density[i];
  }
  return 0;
}

/* ---------------------------------------------------------------- derived material tables */
static void derive_tables(orc_engine *e)
{ /* global_properties.cpp:428-452 (Yeff), 458-483 (Geff), 519-537 (log e), 542-560 (betaeff) */
  for (int i = 1; i <= e->ntypes; i++) for (int j = 1; j <= e->ntypes; j++) {
    const double Yi = e->Y[i], Yj = e->Y[j], vi = e->nu[i], vj = e->nu[j];
    e->Yeff[i][j] = 1. / ((1. - pow(vi, 2.)) / Yi + (1. - pow(vj, 2.)) / Yj);
    e->Geff[i][j] = 1. / (2. * (2. - vi) * (1. + vi) / Yi + 2. * (2. - vj) * (1. + vj) / Yj);
    e->corLog[i][j] = log(e->cor[i][j]);
    e->betaeff[i][j] = e->corLog[i][j] / sqrt(pow(e->corLog[i][j], 2.) + pow(M_PI, 2.));
  }
}

/* ---------------------------------------------------------------- contact model chain
 * sidata-like scratch; follows contact_interface.h:62-186 */
typedef struct {
  int is_wall, itype, jtype, shearupdate;
  double radi, radj, radsum, r, rinv, en[3], delta[3], deltan, meff, mi, mj;
  const double *vi, *vj, *wi, *wj;
  double kn, kt, gamman, gammat, Fn, vn, cri, crj, wr1, wr2, wr3, vtr1, vtr2, vtr3;
  double *hist; int *flag;
  double Fi[3], Ti[3], Fj[3], Tj[3];
} sid_t;

static void surface_default(sid_t *s)
{ /* surface_model_default.h:146-211 */
  const double enx = s->en[0], eny = s->en[1], enz = s->en[2];
  const double vr1 = s->vi[0] - s->vj[0], vr2 = s->vi[1] - s->vj[1], vr3 = s->vi[2] - s->vj[2];
  const double vn = vr1 * enx + vr2 * eny + vr3 * enz;
  const double vn1 = vn * enx, vn2 = vn * eny, vn3 = vn * enz;
  const double vt1 = vr1 - vn1, vt2 = vr2 - vn2, vt3 = vr3 - vn3;
  const double deltan = s->radsum - s->r;
  const double dx = s->delta[0], dy = s->delta[1], dz = s->delta[2], rinv = s->rinv;
  double wr1, wr2, wr3;
  if (s->is_wall) {
    const double cr = s->radi - 0.5 * s->deltan; /* uses the deltan handed in by the wall driver */
    wr1 = cr * s->wi[0] * rinv; wr2 = cr * s->wi[1] * rinv; wr3 = cr * s->wi[2] * rinv;
    s->cri = cr;
  } else {
    const double cri = s->radi - 0.5 * deltan, crj = s->radj - 0.5 * deltan;
    wr1 = (cri * s->wi[0] + crj * s->wj[0]) * rinv;
    wr2 = (cri * s->wi[1] + crj * s->wj[1]) * rinv;
    wr3 = (cri * s->wi[2] + crj * s->wj[2]) * rinv;
    s->cri = cri; s->crj = crj;
  }
  s->vtr1 = vt1 - (dz * wr2 - dy * wr3);
  s->vtr2 = vt2 - (dx * wr3 - dz * wr1);
  s->vtr3 = vt3 - (dy * wr1 - dx * wr2);
  s->vn = vn; s->deltan = deltan; s->wr1 = wr1; s->wr2 = wr2; s->wr3 = wr3;
}

static void normal_apply(sid_t *s, double Fn)
{ /* tail of normal_model_hertz.h:366-383 / normal_model_hooke.h (same code) */
  if (s->is_wall) { for (int d = 0; d < 3; d++) s->Fi[d] += Fn * 1.0 * s->en[d]; }
  else for (int d = 0; d < 3; d++) { s->Fi[d] += Fn * s->en[d]; s->Fj[d] += -s->Fi[d]; }
}

static void normal_hertz(const orc_engine *e, const model_t *m, sid_t *s)
{ /* normal_model_hertz.h:205-266 */
  if (s->flag) *s->flag |= CONTACT_NORMAL;
  const int it = s->itype, jt = s->jtype;
  const double reff = s->is_wall ? s->radi : (s->radi * s->radj / (s->radi + s->radj));
  const double meff = s->meff;
  const double sqrtval = sqrt(reff * s->deltan);
  const double Sn = 2. * e->Yeff[it][jt] * sqrtval;
  const double St = 8. * e->Geff[it][jt] * sqrtval;
  double kn = 4. / 3. * e->Yeff[it][jt] * sqrtval;
  double kt = St;
  const double sqrtFiveOverSix = 0.91287092917527685576161630466800355658790782499663875;
  const double gamman = -2. * sqrtFiveOverSix * e->betaeff[it][jt] * sqrt(Sn * meff);
  const double gammat = m->tangential_damping ? -2. * sqrtFiveOverSix * e->betaeff[it][jt] * sqrt(St * meff) : 0.0;
  kn /= e->nktv2p; kt /= e->nktv2p;
  const double Fn_damping = -gamman * s->vn;
  const double Fn_contact = kn * s->deltan;
  double Fn = Fn_damping + Fn_contact;
  if (m->limitForce && Fn < 0.0) Fn = 0.0;
  s->Fn = Fn; s->kn = kn; s->kt = kt; s->gamman = gamman; s->gammat = gammat;
  normal_apply(s, Fn);
}

static void normal_hooke(const orc_engine *e, const model_t *m, sid_t *s)
{ /* normal_model_hooke.h:230-300 (viscous off) */
  if (s->flag) *s->flag |= CONTACT_NORMAL;
  const int it = s->itype, jt = s->jtype;
  const double reff = s->is_wall ? s->radi : (s->radi * s->radj / (s->radi + s->radj));
  const double meff = s->meff;
  const double sqrtval = sqrt(reff);
  const double coeffRestLogChosen = e->corLog[it][jt];
  double kn = 16. / 15. * sqrtval * (e->Yeff[it][jt]) * pow(15. * meff * e->charVel * e->charVel / (16. * sqrtval * e->Yeff[it][jt]), 0.2);
  double kt = kn;
  if (m->ktToKn) kt *= 0.285714286;
  const double cSq = coeffRestLogChosen * coeffRestLogChosen;
  const double gamman = sqrt(4. * meff * kn * cSq / (cSq + M_PI * M_PI));
  const double gammat = m->tangential_damping ? gamman : 0.0;
  kn /= e->nktv2p; kt /= e->nktv2p;
  const double Fn_damping = -gamman * s->vn;
  const double Fn_contact = kn * s->deltan;
  double Fn = Fn_damping + Fn_contact;
  if (m->limitForce && Fn < 0.0) Fn = 0.0;
  s->Fn = Fn; s->kn = kn; s->kt = kt; s->gamman = gamman; s->gammat = gammat;
  normal_apply(s, Fn);
}

static void tangential_history(const orc_engine *e, const model_t *m, sid_t *s)
{ /* tangential_model_history.h:136-240, 288-334, 404-426 */
  const double enx = s->en[0], eny = s->en[1], enz = s->en[2];
  if (s->flag) *s->flag |= CONTACT_TANGENTIAL;
  double *shear = &s->hist[m->off_shear];
  if (s->shearupdate) {
    const double dt = e->dt;
    shear[0] += s->vtr1 * dt; shear[1] += s->vtr2 * dt; shear[2] += s->vtr3 * dt;
    double rsht = shear[0] * enx + shear[1] * eny + shear[2] * enz;
    shear[0] -= rsht * enx; shear[1] -= rsht * eny; shear[2] -= rsht * enz;
  }
  const double shrmag = sqrt(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2]);
  const double kt = s->kt;
  const double xmu = e->mu[s->itype][s->jtype];
  double Ft1 = -(kt * shear[0]), Ft2 = -(kt * shear[1]), Ft3 = -(kt * shear[2]);
  const double Ft_shear = kt * shrmag;
  const double Ft_friction = xmu * fabs(s->Fn);
  if (Ft_shear > Ft_friction) {
    if (shrmag != 0.0) {
      const double ratio = Ft_friction / Ft_shear;
      Ft1 *= ratio; Ft2 *= ratio; Ft3 *= ratio;
      if (s->shearupdate) { shear[0] = -Ft1 / kt; shear[1] = -Ft2 / kt; shear[2] = -Ft3 / kt; }
    } else Ft1 = Ft2 = Ft3 = 0.0;
  } else {
    const double gammat = s->gammat;
    Ft1 -= (gammat * s->vtr1); Ft2 -= (gammat * s->vtr2); Ft3 -= (gammat * s->vtr3);
  }
  const double tor1 = eny * Ft3 - enz * Ft2, tor2 = enz * Ft1 - enx * Ft3, tor3 = enx * Ft2 - eny * Ft1;
  const double Tn_shear = 0.;
  double ti[3], tj[3] = {0., 0., 0.};
  ti[0] = -s->cri * tor1 + Tn_shear * enx; ti[1] = -s->cri * tor2 + Tn_shear * eny; ti[2] = -s->cri * tor3 + Tn_shear * enz;
  if (!s->is_wall) { tj[0] = -s->crj * tor1 - Tn_shear * enx; tj[1] = -s->crj * tor2 - Tn_shear * eny; tj[2] = -s->crj * tor3 - Tn_shear * enz; }
  if (s->is_wall) {
    const double ar = 1.0;
    s->Fi[0] += Ft1 * ar; s->Fi[1] += Ft2 * ar; s->Fi[2] += Ft3 * ar;
    s->Ti[0] += ti[0] * ar; s->Ti[1] += ti[1] * ar; s->Ti[2] += ti[2] * ar;
  } else {
    s->Fi[0] += Ft1; s->Fi[1] += Ft2; s->Fi[2] += Ft3;
    s->Fj[0] += -Ft1; s->Fj[1] += -Ft2; s->Fj[2] += -Ft3;
    for (int d = 0; d < 3; d++) { s->Ti[d] += ti[d]; s->Tj[d] += tj[d]; }
  }
}

static void rolling_cdt(const orc_engine *e, const model_t *m, sid_t *s)
{ /* rolling_model_cdt.h:91-167 */
  const double rmu = e->rmu[s->itype][s->jtype];
  double rt[3] = {0., 0., 0.};
  const double reff = s->is_wall ? s->radi : (s->radi * s->radj / (s->radi + s->radj));
  const double enx = s->en[0], eny = s->en[1], enz = s->en[2];
  if (s->is_wall) {
    const double wr1 = s->wr1, wr2 = s->wr2, wr3 = s->wr3;
    const double wrmag = sqrt(wr1 * wr1 + wr2 * wr2 + wr3 * wr3);
    if (wrmag > 0.) {
      const double Fn = s->deltan * s->kn;
      rt[0] = rmu * Fn * wr1 / wrmag * reff; rt[1] = rmu * Fn * wr2 / wrmag * reff; rt[2] = rmu * Fn * wr3 / wrmag * reff;
      if (!m->torsionTorque) {
        double dot = rt[0] * enx + rt[1] * eny + rt[2] * enz;
        rt[0] -= enx * dot; rt[1] -= eny * dot; rt[2] -= enz * dot;
      }
    }
  } else {
    double wr[3] = {s->wi[0] - s->wj[0], s->wi[1] - s->wj[1], s->wi[2] - s->wj[2]};
    const double mag = sqrt(wr[0] * wr[0] + wr[1] * wr[1] + wr[2] * wr[2]);
    if (mag > 0.) {
      const double sc = rmu * s->kn * s->deltan * reff / mag;
      rt[0] = wr[0] * sc; rt[1] = wr[1] * sc; rt[2] = wr[2] * sc;
      if (!m->torsionTorque) {
        const double dot = rt[0] * enx + rt[1] * eny + rt[2] * enz;
        rt[0] -= enx * dot; rt[1] -= eny * dot; rt[2] -= enz * dot;
      }
    }
  }
  for (int d = 0; d < 3; d++) { s->Ti[d] -= rt[d]; s->Tj[d] += rt[d]; }
}

static void rolling_epsd(const orc_engine *e, const model_t *m, sid_t *s)
{ /* rolling_model_epsd.h:97-340 ; epsd2 (rolling_model_epsd2.h:152-205): spring kr = kt*reff^2, no dashpot */
  if (s->flag) *s->flag |= CONTACT_ROLLING;
  const double radi = s->radi, radj = s->radj;
  const double reff = s->is_wall ? radi : (radi * radj / (radi + radj));
  double wr1, wr2, wr3, r_inertia;
  if (s->is_wall) {
    wr1 = s->wr1; wr2 = s->wr2; wr3 = s->wr3;
    r_inertia = 1.4 * s->mi * reff * reff;
  } else {
    wr1 = s->wi[0] - s->wj[0]; wr2 = s->wi[1] - s->wj[1]; wr3 = s->wi[2] - s->wj[2];
    const double ri = s->mi * radi * radi, rj = s->mj * radj * radj;
    r_inertia = 1.4 * ri * rj / (ri + rj);
  }
  /* calcRollTorque, rolling_model_epsd.h:262-337 */
  const double enx = s->en[0], eny = s->en[1], enz = s->en[2], dt = e->dt;
  double *ch = &s->hist[m->off_roll];
  const double rmu = e->rmu[s->itype][s->jtype];
  double wt[3];
  if (m->torsionTorque) { wt[0] = wr1; wt[1] = wr2; wt[2] = wr3; }
  else { const double dot = wr1 * enx + wr2 * eny + wr3 * enz; wt[0] = wr1 - enx * dot; wt[1] = wr2 - eny * dot; wt[2] = wr3 - enz * dot; }
  const double kr = (m->rolling == R_EPSD2) ? s->kt * reff * reff : 2.25 * s->kn * rmu * rmu * reff * reff;
  double rt[3];
  for (int d = 0; d < 3; d++) rt[d] = ch[d] + wt[d] * (dt * kr);
  const double mag = sqrt(rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2]);
  const double tmax = fabs(s->Fn) * reff * rmu;
  if (mag > tmax) {
    const double factor = tmax / mag;
    for (int d = 0; d < 3; d++) rt[d] *= factor;
    if (s->shearupdate) for (int d = 0; d < 3; d++) ch[d] = rt[d];
  } else {
    if (s->shearupdate) for (int d = 0; d < 3; d++) ch[d] = rt[d];
    if (m->rolling == R_EPSD) {
      const double r_coef = e->rvisc[s->itype][s->jtype] * 2 * sqrt(r_inertia * kr);
      for (int d = 0; d < 3; d++) rt[d] += r_coef * wt[d];
    }
  }
  for (int d = 0; d < 3; d++) { s->Ti[d] -= rt[d]; s->Tj[d] += rt[d]; }
}

/* ContactModel::surfacesIntersect, contact_models.h:228-238 */
static void chain_intersect(const orc_engine *e, const model_t *m, sid_t *s)
{
  surface_default(s);
  if (m->normal == N_HERTZ) normal_hertz(e, m, s); else normal_hooke(e, m, s);
  if (m->tangential) tangential_history(e, m, s);
  if (m->rolling == R_CDT) rolling_cdt(e, m, s);
  else if (m->rolling == R_EPSD || m->rolling == R_EPSD2) rolling_epsd(e, m, s);
}
/* ContactModel::surfacesClose, contact_models.h:246-253 */
static void chain_close(const model_t *m, double *hist, int *flag)
{
  if (m->tangential) { if (flag) *flag &= ~CONTACT_TANGENTIAL; for (int d = 0; d < 3; d++) hist[m->off_shear + d] = 0.0; }
  if (m->off_roll >= 0) { if (flag) *flag &= ~CONTACT_ROLLING; for (int d = 0; d < 3; d++) hist[m->off_roll + d] = 0.0; }
}

/* ---------------------------------------------------------------- neighbour build */
static void pbc_wrap(orc_engine *e)
{ /* domain.cpp Domain::pbc(): owned particles re-enter a periodic box */
  for (long i = 0; i < e->n; i++) for (int d = 0; d < 3; d++) if (e->periodic[d]) {
    double *xx = &e->x[3 * i + d];
    if (*xx < e->lo[d]) *xx += e->prd[d];
    if (*xx >= e->hi[d]) { *xx -= e->prd[d]; if (*xx < e->lo[d]) *xx = e->lo[d]; } /* domain.cpp: x = MAX(x,lo) after the hi wrap */
  }
}

static void build(orc_engine *e)
{ /* neigh_gran.cpp:485-644 (granular_bin_no_newton) incl. history remap :590-625 */
  const long n = e->n; const int dnum = e->pm.dnum;
  pbc_wrap(e);
  double rmax = 0.0; for (long i = 0; i < n; i++) if (e->radius[i] > rmax) rmax = e->radius[i];
  const double cutmax = 2.0 * rmax * e->cdf + e->skin; /* pair_gran.cpp:591-603 + skin */
  /* bins */
  int nb[3]; double binsz[3], blo[3];
  for (int d = 0; d < 3; d++) {
    double lo = e->lo[d], hi = e->hi[d];
    if (!e->periodic[d]) for (long i = 0; i < n; i++) { double xx = e->x[3 * i + d]; if (xx < lo) lo = xx; if (xx > hi) hi = xx; }
    nb[d] = (int)floor((hi - lo) / cutmax); if (nb[d] < 1) nb[d] = 1; if (nb[d] > 512) nb[d] = 512;
    binsz[d] = (hi - lo) / nb[d]; blo[d] = lo;
  }
  const long nbins = (long)nb[0] * nb[1] * nb[2];
  int *head = (int *)malloc(sizeof(int) * nbins), *next = (int *)malloc(sizeof(int) * (n ? n : 1));
  for (long b = 0; b < nbins; b++) head[b] = -1;
  int *bin3 = (int *)malloc(sizeof(int) * 3 * (n ? n : 1));
  for (long i = n - 1; i >= 0; i--) {
    long b = 0; int c[3];
    for (int d = 0; d < 3; d++) { c[d] = (int)floor((e->x[3 * i + d] - blo[d]) / binsz[d]); if (c[d] < 0) c[d] = 0; if (c[d] >= nb[d]) c[d] = nb[d] - 1; bin3[3 * i + d] = c[d]; }
    b = ((long)c[2] * nb[1] + c[1]) * nb[0] + c[0];
    next[i] = head[b]; head[b] = (int)i;
  }
  /* old list -> used for the history lookup */
  long *ofirst = e->first; int *onum = e->numneigh, *ojl = e->jlist, *oflag = e->flag; double *ohist = e->hist; signed char *oshift = e->jshift;
  long cap = e->cap > 0 ? e->cap : 16 * n + 64;
  long *first = (long *)malloc(sizeof(long) * (n + 1)); int *num = (int *)calloc(n ? n : 1, sizeof(int));
  int *jl = (int *)malloc(sizeof(int) * cap), *fl = (int *)malloc(sizeof(int) * cap); signed char *sh = (signed char *)malloc(3 * cap);
  double *hs = (double *)malloc(sizeof(double) * cap * (dnum ? dnum : 1));
  long np = 0;
  for (long i = 0; i < n; i++) {
    first[i] = np;
    const double xi = e->x[3 * i], yi = e->x[3 * i + 1], zi = e->x[3 * i + 2], radi = e->radius[i];
    for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
      int c[3] = {bin3[3 * i] + dx, bin3[3 * i + 1] + dy, bin3[3 * i + 2] + dz}, s[3] = {0, 0, 0}, ok = 1;
      for (int d = 0; d < 3; d++) {
        if (c[d] < 0) { if (e->periodic[d]) { c[d] += nb[d]; s[d] = -1; } else ok = 0; }
        else if (c[d] >= nb[d]) { if (e->periodic[d]) { c[d] -= nb[d]; s[d] = 1; } else ok = 0; }
      }
      if (!ok) continue;
      for (int j = head[((long)c[2] * nb[1] + c[1]) * nb[0] + c[0]]; j >= 0; j = next[j]) {
        if (j <= i) continue; /* :576 */
        /* image of j as the reference's ghost would carry it: x_j + s*prd */
        const double xj = s[0] ? e->x[3 * j] + s[0] * e->prd[0] : e->x[3 * j];
        const double yj = s[1] ? e->x[3 * j + 1] + s[1] * e->prd[1] : e->x[3 * j + 1];
        const double zj = s[2] ? e->x[3 * j + 2] + s[2] * e->prd[2] : e->x[3 * j + 2];
        const double delx = xi - xj, dely = yi - yj, delz = zi - zj;
        const double rsq = delx * delx + dely * dely + delz * delz;
        const double radsum = (radi + e->radius[j]) * e->cdf;
        const double cutsq = (radsum + e->skin) * (radsum + e->skin);
        if (rsq <= cutsq) { /* :587 */
          if (np == cap) { cap *= 2; jl = realloc(jl, sizeof(int) * cap); fl = realloc(fl, sizeof(int) * cap); sh = realloc(sh, 3 * cap); hs = realloc(hs, sizeof(double) * cap * (dnum ? dnum : 1)); }
          jl[np] = j; sh[3 * np] = (signed char)s[0]; sh[3 * np + 1] = (signed char)s[1]; sh[3 * np + 2] = (signed char)s[2];
          fl[np] = 0; for (int d = 0; d < dnum; d++) hs[np * dnum + d] = 0.0;
          if (dnum && ofirst && rsq < radsum * radsum) { /* :592 */
            for (long m = ofirst[i]; m < ofirst[i] + onum[i]; m++)
              if (ojl[m] == j && oflag[m]) { /* partner found (only flagged pairs are partners, fix_contact_history.cpp:351) */
                fl[np] = 1; for (int d = 0; d < dnum; d++) hs[np * dnum + d] = ohist[m * dnum + d]; break;
              }
          }
          np++; num[i]++;
        }
      }
    }
  }
  first[n] = np;
  free(ofirst); free(onum); free(ojl); free(oflag); free(ohist); free(oshift);
  e->first = first; e->numneigh = num; e->jlist = jl; e->flag = fl; e->hist = hs; e->jshift = sh; e->npairs = np; e->cap = cap;
  free(head); free(next); free(bin3);
  memcpy(e->xhold, e->x, sizeof(double) * 3 * n); /* neighbor.cpp:1486-1510 */
  /* primitive wall candidate lists: FixWallGran::pre_force fix_wall_gran.cpp:688-710, primitive_wall.h:129-138 */
  for (int w = 0; w < e->nwalls; w++) {
    wall_t *W = &e->walls[w]; W->ncand = 0;
    if (!W->cand) W->cand = (int *)malloc(sizeof(int) * (n ? n : 1));
    for (long i = 0; i < n; i++) {
      int in;
      if (W->wtype < 3) { /* Plane::resolveNeighlist primitive_wall_definitions.h:144-150 */
        double dMax = e->radius[i] + e->skin, dist = e->x[3 * i + W->wtype] - W->param[0];
        double absdist = (dist > 0.0) ? dist : -dist; in = (absdist <= dMax);
      } else { /* Cylinder::resolveNeighlist :195-201 */
        const int dd = W->wtype - 3; double dy = e->x[3 * i + (dd + 1) % 3] - W->param[1], dz = e->x[3 * i + (dd + 2) % 3] - W->param[2];
        double dMax = e->radius[i] + e->skin, dist = sqrt(dy * dy + dz * dz) - W->param[0];
        in = (dMax < dist || -dMax < dist);
      }
      if (in) W->cand[W->ncand++] = (int)i;
    }
  }
  e->nbuilds++; e->ago = 0;
}

/* ---------------------------------------------------------------- forces */
static void force_clear(orc_engine *e) { memset(e->f, 0, sizeof(double) * 3 * e->n); memset(e->torque, 0, sizeof(double) * 3 * e->n); }

static void pair_compute(orc_engine *e, int shearupdate)
{ /* pair_gran_base.h:187-508 */
  const int dnum = e->pm.dnum; const double cdm = e->cdf * e->cdf;
  for (long i = 0; i < e->n; i++) {
    const double xtmp = e->x[3 * i], ytmp = e->x[3 * i + 1], ztmp = e->x[3 * i + 2], radi = e->radius[i];
    for (long m = e->first[i]; m < e->first[i] + e->numneigh[i]; m++) {
      const int j = e->jlist[m]; const signed char *s = &e->jshift[3 * m];
      const double xj = s[0] ? e->x[3 * j] + s[0] * e->prd[0] : e->x[3 * j];
      const double yj = s[1] ? e->x[3 * j + 1] + s[1] * e->prd[1] : e->x[3 * j + 1];
      const double zj = s[2] ? e->x[3 * j + 2] + s[2] * e->prd[2] : e->x[3 * j + 2];
      const double delx = xtmp - xj, dely = ytmp - yj, delz = ztmp - zj;
      const double rsq = delx * delx + dely * dely + delz * delz;
      const double radj = e->radius[j], radsum = radi + radj;
      if (rsq < radsum * radsum) { /* :358 */
        sid_t s_; sid_t *sd = &s_; memset(sd, 0, sizeof *sd);
        const double r = sqrt(rsq), rinv = 1.0 / r;
        double mi = e->rmass[i], mj = e->rmass[j];
        double meff = mi * mj / (mi + mj);
        if (e->mask[i] & e->freezebit) meff = mj; /* :389-393 */
        if (e->mask[j] & e->freezebit) meff = mi;
        sd->is_wall = 0; sd->itype = e->type[i]; sd->jtype = e->type[j]; sd->shearupdate = shearupdate;
        sd->radi = radi; sd->radj = radj; sd->radsum = radsum; sd->r = r; sd->rinv = rinv;
        sd->delta[0] = delx; sd->delta[1] = dely; sd->delta[2] = delz;
        sd->en[0] = delx * rinv; sd->en[1] = dely * rinv; sd->en[2] = delz * rinv;
        sd->meff = meff; sd->mi = mi; sd->mj = mj;
        sd->vi = &e->v[3 * i]; sd->vj = &e->v[3 * j]; sd->wi = &e->omega[3 * i]; sd->wj = &e->omega[3 * j];
        sd->hist = dnum ? &e->hist[m * dnum] : NULL; sd->flag = &e->flag[m];
        chain_intersect(e, &e->pm, sd);
        for (int d = 0; d < 3; d++) { e->f[3 * i + d] += sd->Fi[d]; e->torque[3 * i + d] += sd->Ti[d]; e->f[3 * j + d] += sd->Fj[d]; e->torque[3 * j + d] += sd->Tj[d]; }
      } else if (rsq < cdm * radsum * radsum) { /* :420 ; unreachable when cdf == 1 */
        chain_close(&e->pm, dnum ? &e->hist[m * dnum] : NULL, &e->flag[m]);
      }
    }
  }
}

static void wall_compute(orc_engine *e, wall_t *W, int shearupdate)
{ /* fix_wall_gran.cpp:988-1121 + fix_wall_gran_base.h:159-367 */
  const int dnum = W->m.dnum; const double cdmul = e->cdf - 1.0;
  double rmax = 0.0; for (long i = 0; i < e->n; i++) if (e->radius[i] > rmax) rmax = e->radius[i];
  const double cutneighmax = 2.0 * rmax * e->cdf + e->skin;
  for (int c = 0; c < W->ncand; c++) {
    const int ip = W->cand[c]; const double *pos = &e->x[3 * ip]; const double r = e->radius[ip];
    double delta[3] = {0, 0, 0}, deltan, v_wall[3] = {0., 0., 0.};
    if (W->shear) v_wall[W->shearDim] = W->vshear;
    if (W->wtype < 3) { /* Plane::resolveContact primitive_wall_definitions.h:133-142 */
      const int dx = W->wtype; const double p = W->param[0];
      delta[dx] = p - pos[dx];
      deltan = pos[dx] > p ? pos[dx] - p - r : p - pos[dx] - r;
    } else { /* Cylinder::resolveContact :170-193 */
      const int dd = W->wtype - 3, iy = (dd + 1) % 3, iz = (dd + 2) % 3; const double R = W->param[0];
      const double dy = pos[iy] - W->param[1], dz = pos[iz] - W->param[2], dist = sqrt(dy * dy + dz * dz);
      if (dist == 0.0) { deltan = 0.0; }
      else if (dist > R) { deltan = dist - R - r; const double fact = (dist - R) / dist; delta[iy] = -dy * fact; delta[iz] = -dz * fact; }
      else { deltan = R - dist - r; const double fact = (R - dist) / dist; delta[iy] = dy * fact; delta[iz] = dz * fact; }
    }
    double *hist = dnum ? &W->hist[(long)ip * dnum] : NULL;
    if (deltan > cutneighmax) continue;
    if (deltan <= 0 || deltan < cdmul * r) {
      const int intersect = (deltan <= 0);
      if (W->shear && W->shearAxis >= 0) { /* calcRadialDistance + cross(shearAxisVec, rdist) fix_wall_gran.cpp:1080-1084 */
        const int dd = W->wtype - 3; double rd[3] = {0, 0, 0};
        rd[(dd + 1) % 3] = pos[(dd + 1) % 3] - W->param[1]; rd[(dd + 2) % 3] = pos[(dd + 2) % 3] - W->param[2];
        const double *a = W->shearAxisVec;
        v_wall[0] = a[1] * rd[2] - a[2] * rd[1]; v_wall[1] = a[2] * rd[0] - a[0] * rd[2]; v_wall[2] = a[0] * rd[1] - a[1] * rd[0];
      }
      if (intersect) {
        sid_t s_; sid_t *sd = &s_; memset(sd, 0, sizeof *sd);
        sd->is_wall = 1; sd->radi = r; sd->deltan = -deltan;
        sd->delta[0] = -delta[0]; sd->delta[1] = -delta[1]; sd->delta[2] = -delta[2];
        sd->vi = &e->v[3 * ip]; sd->vj = v_wall; sd->wi = &e->omega[3 * ip]; sd->wj = NULL;
        sd->r = sd->radi - sd->deltan; /* fix_wall_gran_base.h:194 */
        sd->rinv = 1.0 / sd->r;
        sd->itype = e->type[ip]; sd->jtype = W->atom_type; sd->meff = e->rmass[ip]; sd->mi = e->rmass[ip];
        sd->shearupdate = shearupdate; sd->radsum = sd->radi;
        for (int d = 0; d < 3; d++) sd->en[d] = sd->delta[d] * sd->rinv;
        sd->hist = hist; sd->flag = NULL;
        chain_intersect(e, &W->m, sd);
        for (int d = 0; d < 3; d++) { e->f[3 * ip + d] += sd->Fi[d]; e->torque[3 * ip + d] += sd->Ti[d]; }
      } else chain_close(&W->m, hist, NULL);
    } else if (hist) for (int d = 0; d < dnum; d++) hist[d] = 0.0; /* :1117-1119 */
  }
}

static void compute_forces(orc_engine *e, int shearupdate)
{ /* verlet.cpp:337-369: force_clear, pair, post_force fixes (gravity, walls, freeze) */
  force_clear(e);
  if (e->have_pair) pair_compute(e, shearupdate);
  if (e->have_gravity) for (long i = 0; i < e->n; i++) if (e->mask[i] & 1) { /* fix_gravity.cpp:331-339, group all */
    const double m = e->rmass[i]; e->f[3 * i] += m * e->g[0]; e->f[3 * i + 1] += m * e->g[1]; e->f[3 * i + 2] += m * e->g[2]; }
  for (int w = 0; w < e->nwalls; w++) wall_compute(e, &e->walls[w], shearupdate);
  if (e->freezebit) for (long i = 0; i < e->n; i++) if (e->mask[i] & e->freezebit) for (int d = 0; d < 3; d++) { e->f[3 * i + d] = 0.0; e->torque[3 * i + d] = 0.0; } /* fix_freeze.cpp:132-144 */
}

int orc_setup(orc_engine *e)
{ /* Verlet::setup verlet.cpp:134-199 */
  if (!e->n && !e->tag) return fail(e, "no particles uploaded");
  derive_tables(e);
  for (int w = 0; w < e->nwalls; w++) if (!e->walls[w].hist) { e->walls[w].hist = (double *)calloc((size_t)(e->n ? e->n : 1) * (e->walls[w].m.dnum ? e->walls[w].m.dnum : 1), sizeof(double)); }
  build(e);
  e->nbuilds = 0; /* neighbor->ncalls counts builds of the current run only (neighbor.cpp init: ncalls = 0) */
  compute_forces(e, 0);
  e->setup_done = 1; return 0;
}

int orc_run(orc_engine *e, long nsteps)
{ /* Verlet::run verlet.cpp:264-391 */
  if (!e->setup_done) return fail(e, "run before setup");
  const double dtv = e->dt, dtf = 0.5 * e->dt * e->ftm2v, dtfrotate = dtf / 0.4; /* fix_nve.cpp:86, fix_nve_sphere.cpp:69,150 */
  for (long s = 0; s < nsteps; s++) {
    e->ntimestep++;
    for (long i = 0; i < e->n; i++) if (e->mask[i] & e->integbit) { /* fix_nve_sphere.cpp:134-183 */
      const double dtfm = dtf / (e->rmass[i] * (1. + 0.0 / e->density[i]));
      for (int d = 0; d < 3; d++) { e->v[3 * i + d] += dtfm * e->f[3 * i + d]; e->x[3 * i + d] += dtv * e->v[3 * i + d]; }
      const double dtirotate = dtfrotate / (e->radius[i] * e->radius[i] * e->rmass[i]);
      for (int d = 0; d < 3; d++) e->omega[3 * i + d] += dtirotate * e->torque[3 * i + d];
    }
    /* Neighbor::decide neighbor.cpp:1362-1376 + check_distance :1425-1466 */
    int nflag = 0; e->ago++;
    if (e->ago >= e->delay && e->ago % e->every == 0) {
      if (!e->check) nflag = 1;
      else { const double deltasq = 0.25 * e->skin * e->skin;
        for (long i = 0; i < e->n; i++) { const double dx = e->x[3 * i] - e->xhold[3 * i], dy = e->x[3 * i + 1] - e->xhold[3 * i + 1], dz = e->x[3 * i + 2] - e->xhold[3 * i + 2];
          if (dx * dx + dy * dy + dz * dz > deltasq) nflag = 1; } }
    }
    if (nflag) build(e);
    compute_forces(e, 1);
    for (long i = 0; i < e->n; i++) if (e->mask[i] & e->integbit) { /* fix_nve_sphere.cpp:205-244 */
      const double dtfm = dtf / (e->rmass[i] * (1. + 0.0 / e->density[i]));
      for (int d = 0; d < 3; d++) e->v[3 * i + d] += dtfm * e->f[3 * i + d];
      const double dtirotate = dtfrotate / (e->radius[i] * e->radius[i] * e->rmass[i]);
      for (int d = 0; d < 3; d++) e->omega[3 * i + d] += dtirotate * e->torque[3 * i + d];
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- read-back (by ascending tag) */
static int cmp_tag(const void *a, const void *b) { const long *x = a, *y = b; return (x[0] > y[0]) - (x[0] < y[0]); }
static long *tag_order(const orc_engine *e)
{ long *o = (long *)malloc(sizeof(long) * 2 * (e->n ? e->n : 1)); for (long i = 0; i < e->n; i++) { o[2 * i] = e->tag[i]; o[2 * i + 1] = i; } qsort(o, e->n, 2 * sizeof(long), cmp_tag); return o; }

long orc_nlocal(const orc_engine *e) { return e->n; }
int orc_download(orc_engine *e, const char *field, void *out, long count)
{
  if (count != e->n) return fail(e, "count != nlocal");
  long *o = tag_order(e); int rc = 0;
  const int *isrc = !strcmp(field, "tag") ? e->tag : !strcmp(field, "type") ? e->type : !strcmp(field, "mask") ? e->mask : NULL;
  const double *s1 = !strcmp(field, "radius") ? e->radius : !strcmp(field, "rmass") ? e->rmass : !strcmp(field, "density") ? e->density : NULL;
  const double *s3 = !strcmp(field, "x") ? e->x : !strcmp(field, "v") ? e->v : !strcmp(field, "f") ? e->f : !strcmp(field, "omega") ? e->omega : !strcmp(field, "torque") ? e->torque : NULL;
  if (isrc) for (long k = 0; k < e->n; k++) ((int *)out)[k] = isrc[o[2 * k + 1]];
  else if (s1) for (long k = 0; k < e->n; k++) ((double *)out)[k] = s1[o[2 * k + 1]];
  else if (s3) for (long k = 0; k < e->n; k++) for (int d = 0; d < 3; d++) ((double *)out)[3 * k + d] = s3[3 * o[2 * k + 1] + d];
  else rc = fail(e, "unknown field");
  free(o); return rc;
}
int orc_pair_count(orc_engine *e, long *np, int *dnum) { *np = e->npairs; *dnum = e->pm.dnum; return 0; }
typedef struct { int lo, hi, flag; long m; int swap; } prow_t;
static int cmp_prow(const void *a, const void *b) { const prow_t *x = a, *y = b; if (x->lo != y->lo) return (x->lo > y->lo) - (x->lo < y->lo); return (x->hi > y->hi) - (x->hi < y->hi); }
int orc_download_pairs(orc_engine *e, int *lo, int *hi, int *flag, double *hist)
{
  const int dnum = e->pm.dnum; prow_t *rows = (prow_t *)malloc(sizeof(prow_t) * (e->npairs ? e->npairs : 1)); long k = 0;
  for (long i = 0; i < e->n; i++) for (long m = e->first[i]; m < e->first[i] + e->numneigh[i]; m++) {
    const int ti = e->tag[i], tj = e->tag[e->jlist[m]]; rows[k].swap = ti > tj; rows[k].lo = ti < tj ? ti : tj; rows[k].hi = ti < tj ? tj : ti; rows[k].flag = e->flag[m]; rows[k].m = m; k++; }
  qsort(rows, k, sizeof(prow_t), cmp_prow);
  for (long r = 0; r < k; r++) { lo[r] = rows[r].lo; hi[r] = rows[r].hi; if (flag) flag[r] = rows[r].flag;
    if (hist) for (int d = 0; d < dnum; d++) hist[r * dnum + d] = rows[r].swap ? -e->hist[rows[r].m * dnum + d] : e->hist[rows[r].m * dnum + d]; }
  free(rows); return 0;
}
int orc_download_wall_history(orc_engine *e, const char *id, double *out, long count)
{
  if (count != e->n) return fail(e, "count != nlocal");
  for (int w = 0; w < e->nwalls; w++) if (!strcmp(e->walls[w].id, id)) {
    const int dnum = e->walls[w].m.dnum; long *o = tag_order(e);
    for (long k = 0; k < e->n; k++) for (int d = 0; d < dnum; d++) out[k * dnum + d] = e->walls[w].hist[o[2 * k + 1] * dnum + d];
    free(o); return 0; }
  return fail(e, "no such wall");
}
typedef struct { long ntimestep, nbuilds, nlocal, nghost, npairs_full, ncontacts_full, kernel_launches; int maxneigh, dnum; double step_kernel_ms; long step_kernel_calls; } orc_stats;
int orc_get_stats(orc_engine *e, orc_stats *s)
{ memset(s, 0, sizeof *s); s->ntimestep = e->ntimestep; s->nbuilds = e->nbuilds; s->nlocal = e->n; s->npairs_full = 2 * e->npairs; s->dnum = e->pm.dnum;
  long c = 0; for (long m = 0; m < e->npairs; m++) c += e->flag[m] != 0; s->ncontacts_full = 2 * c; return 0; }
