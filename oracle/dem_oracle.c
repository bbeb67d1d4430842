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

enum { N_HERTZ = 1, N_HOOKE = 2, N_HYST1 = 3, N_HYST2 = 4 };
enum { R_OFF = 0, R_CDT = 1, R_EPSD = 2, R_EPSD2 = 3 };
enum { CONTACT_NORMAL = 1, CONTACT_COHESION = 2, CONTACT_TANGENTIAL = 4, CONTACT_ROLLING = 8 }; /* contact_model_constants.h:64-69 (values irrelevant: only != 0 is tested) */

enum { C_OFF = 0, C_BOND = 1, C_BONDNL = 2 };
typedef struct {
  int normal, tangential, rolling;
  int tangential_damping, limitForce, torsionTorque, ktToKn;
  int cdtnl2; /* rolling_friction cdtnonlinear2 */
  int dissipation; /* cohesion bond: dissipationBond on (cohesion_model_bond.h:270) */
  int dnum, off_shear, off_roll, off_norm; /* off_norm: the 12 history values of normal model hysteretic/nonlinear1|2 */
  /* cohesion bond / bond/nonlinear: cohesion_model_bond.h:254-276, cohesion_model_bond_nonlinear.h:233-246 */
  int cohesion, off_bond;
  int stressBreak, tension, compression, shear, ntorque, ttorque, createAlways, damping, dampingSmooth, ratioTC;
  signed char nflag[40]; /* newtonflag of every history value (1: flips sign with the pair orientation) */
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

#define MAXMESH 8
#define NUM_NEIGH_MAX 5      /* tri_mesh.h:67 SurfaceMesh<3,5> */
#define EPSILON_PRECISION 1e-8   /* multi_node_mesh.h:59 */
#define EPSILON_CURVATURE 0.00001 /* surface_mesh.h:61 */
#define SMALL_TRIMESH (1.e-10)   /* tri_mesh_I.h:47-50 */
#define LARGE_TRIMESH 1000000
#define SMALL_DELTA_MESH 1e-6

typedef struct {  /* TriMesh + FixNeighlistMesh + FixContactHistoryMesh of one `fix mesh/surface` */
  char id[64]; int atom_type; int ntri; int wall; /* wall: index of the mesh wall fix using it (-1 none) */
  double (*node)[3][3], (*center)[3], *rbound, (*edgeVec)[3][3], (*edgeLen)[3], (*surfNorm)[3], (*edgeNorm)[3][3];
  int *obtuse, *nNeighs, (*neighFaces)[NUM_NEIGH_MAX]; unsigned char (*edgeActive)[3], (*cornerActive)[3];
  double curvature, precision;
  int stress; double f_total[3], torque_total[3], p_ref[3]; /* fix mesh/surface/stress: mesh_module_stress.cpp:286-345,479-488 */
  int moving; /* 0 static, 1 `linear`, 2 `rotate` */ double vel[3]; double rot_origin[3], rot_axis[3], rot_omega; double (*vnode)[3][3]; double (*nodesLastRe)[3][3]; int next_reneighbor;
  /* FixNeighlistMesh: per-triangle particle lists */
  int **contacts; int *ncontacts, *capcontacts;
  /* FixContactHistoryMesh: per-particle rows sized by the particle's candidate count */
  int *nneighs, *npartner; int **partner; double **chist; unsigned char **keep;
} mesh_t;
typedef struct { char id[64]; model_t m; int nmesh; int mesh[MAXMESH]; } meshwall_t;

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
  /* bond properties (peratomtypepair unless noted), index = enum BP_* */
  double bp[40][MAXT + 1][MAXT + 1]; double tsCreateBond; double rmin;
  model_t pm; int have_pair;
  wall_t walls[MAXW]; int nwalls;
  mesh_t meshes[MAXMESH]; int nmeshes; meshwall_t mwalls[MAXMESH]; int nmwalls;
  double g[3]; int have_gravity;
  int freezebit, integbit;
  double cdf; /* neighbor->contactDistanceFactor (1.0 without bond models) */
  double cdf_user; /* neigh_modify contact_distance_factor (neighbor.cpp:1922-1925), 0 = not given */
  /* particles */
  long n; int *tag, *type, *mask;
  double *x, *v, *f, *omega, *torque, *radius, *rmass, *density, *xhold;
  /* half list (CSR) + history */
  long *first; int *numneigh; int *jlist; signed char *jshift; int *flag; double *hist; long npairs, cap;
  long bond_created, bond_broken; /* compute bond/counter (compute_bond_counter.cpp:140-156), linear bond model only */
  long ntimestep, nbuilds; int ago; int setup_done;
  int nxf; struct { char id[64]; int kind, bit; double v[3]; } xf[4]; /* fix addforce (kind 0) / fix viscous (kind 1), in the order of definition */
  int ins_mass, ins_open; /* orc_insert_step_*: mass as fix insert forms it / the timestep is half done */
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
int orc_set_contact_distance_factor(orc_engine *e, double f) { if (f < 1.0) return fail(e, "Illegal neigh_modify command. Please set contact_distance_factor value >=1"); e->cdf_user = f; return 0; }

/* peratomtypepair properties of the two bond models (cohesion_model_bond.h:76-178, cohesion_model_bond_nonlinear.h:77-147) */
enum { BP_LAMBDA = 0, BP_KN, BP_KT, BP_DFN, BP_DFT, BP_DTN, BP_DTT, BP_MAXDIST, BP_MAXSIGMA, BP_MAXTAU, BP_CREATEDIST, BP_RATIOTC,
       BP_K_FN1, BP_KU_FN1, BP_KC_FN1, BP_K_FN2, BP_KU_FN2, BP_KC_FN2, BP_K_FT, BP_K_TN, BP_KU_TN, BP_KC_TN, BP_K_TT, BP_KU_TT, BP_KC_TT, BP_COUNT,
       /* normal models hysteretic/nonlinear1|2 (normal_model_hysteretic_nonlinear1.h:105-117) */
       HP_KEL = BP_COUNT, HP_KN2K1, HP_KN2KC, HP_PHIF, HP_FADH, HP_ALPHA, HP_CIN, HP_A1, HP_A2, HP_A3, HP_KCIN,
       /* linear bond, option dissipationBond (cohesion_model_bond.h:352-362; stored as given, inverted where used) */
       DP_FN, DP_FT, DP_TN, DP_TT, HP_END };
static int bond_prop_index(const char *name)
{
  static const char *lin[] = {"radiusMultiplierBond", "normalBondStiffnessPerUnitArea", "tangentialBondStiffnessPerUnitArea", "dampingNormalForceBond",
    "dampingTangentialForceBond", "dampingNormalTorqueBond", "dampingTangentialTorqueBond", "maxDistanceBond", "maxSigmaBond", "maxTauBond", "createDistanceBond", "ratioTensionCompression"};
  static const char *nl[] = {"radiusMultiplierBondnonlinear", "", "", "dampingNormalForceBondnonlinear", "dampingTangentialForceBondnonlinear",
    "dampingNormalTorqueBondnonlinear", "dampingTangentialTorqueBondnonlinear", "maxDistanceBondnonlinear", "maxSigmaBondnonlinear", "maxTauBondnonlinear",
    "createDistanceBondnonlinear", "ratioTensionCompressionBondnonlinear", "stiffnessPerUnitAreaK_fn1", "stiffnessPerUnitAreaKu_fn1", "stiffnessPerUnitAreaKc_fn1",
    "stiffnessPerUnitAreaK_fn2", "stiffnessPerUnitAreaKu_fn2", "stiffnessPerUnitAreaKc_fn2", "stiffnessPerUnitAreaK_ft", "stiffnessPerUnitAreaK_tn",
    "stiffnessPerUnitAreaKu_tn", "stiffnessPerUnitAreaKc_tn", "stiffnessPerUnitAreaK_tt", "stiffnessPerUnitAreaKu_tt", "stiffnessPerUnitAreaKc_tt"};
  static const char *hy[] = {"LoadingStiffness", "UnloadingStiffness", "coefficientAdhesionStiffness", "coefficientPlasticityDepth", "pullOffForce",
    "alphaCustom", "cinCustom", "aoneCustom", "atwoCustom", "athreeCustom", "kcinCustom"};
  for (int k = 0; k < 11; k++) if (!strcmp(name, hy[k])) return HP_KEL + k;
  static const char *dp[] = {"dissipationNormalForceBond", "dissipationTangentialForceBond", "dissipationNormalTorqueBond", "dissipationTangentialTorqueBond"};
  for (int k = 0; k < 4; k++) if (!strcmp(name, dp[k])) return DP_FN + k;
  for (int k = 0; k < 12; k++) if (!strcmp(name, lin[k])) return k;
  for (int k = 0; k < BP_COUNT; k++) if (nl[k][0] && !strcmp(name, nl[k])) return k;
  return -1;
}

int orc_set_property(orc_engine *e, const char *name, const char *kind, const double *v, int n)
{
  const int T = e->ntypes;
  if (!strcmp(kind, "scalar")) {
    if (!strcmp(name, "characteristicVelocity")) { e->charVel = v[0]; return 0; }
    if (!strcmp(name, "tsCreateBond") || !strcmp(name, "tsCreateBondnonlinear")) { e->tsCreateBond = v[0]; return 0; }
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
    if (!dst) { const int b = bond_prop_index(name); if (b >= 0) dst = e->bp[b]; }
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
    else if (!strcmp(a[1], "hysteretic/nonlinear1")) { m->normal = N_HYST1; m->limitForce = 1; } /* limitForce defaults to on: :100 */
    else if (!strcmp(a[1], "hysteretic/nonlinear2")) { m->normal = N_HYST2; m->limitForce = 1; }
    else return fail(e, "normal model not supported");
    a += 2; argc -= 2;
  } else return fail(e, "expected 'model'");
  if (argc > 1 && !strcmp(a[0], "tangential")) {
    if (!strcmp(a[1], "history")) m->tangential = 1; else if (!strcmp(a[1], "hysteretic/nonlinear")) m->tangential = 2; else return fail(e, "tangential model not supported");
    a += 2; argc -= 2;
  }
  m->tension = m->compression = m->shear = m->ntorque = m->ttorque = m->damping = 1;
  if (argc > 1 && !strcmp(a[0], "cohesion")) {
    if (!strcmp(a[1], "bond")) m->cohesion = C_BOND; else if (!strcmp(a[1], "bond/nonlinear")) m->cohesion = C_BONDNL;
    else if (!strcmp(a[1], "off")) m->cohesion = C_OFF; else return fail(e, "cohesion model not supported");
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "rolling_friction")) {
    if (!strcmp(a[1], "cdt")) m->rolling = R_CDT; else if (!strcmp(a[1], "cdtnonlinear2")) { m->rolling = R_CDT; m->cdtnl2 = 1; } else if (!strcmp(a[1], "epsd")) m->rolling = R_EPSD;
    else if (!strcmp(a[1], "epsd2")) m->rolling = R_EPSD2; else if (!strcmp(a[1], "off")) m->rolling = R_OFF;
    else return fail(e, "rolling model not supported");
    a += 2; argc -= 2;
  }
  /* history slot order = model construction order: cohesion, tangential, rolling (contact_models.h:141-145) */
  m->dnum = 0; m->off_shear = m->off_roll = m->off_bond = m->off_norm = -1;
  if (m->normal == N_HYST1 || m->normal == N_HYST2) { m->off_norm = m->dnum; m->nflag[m->dnum + 10] = m->nflag[m->dnum + 11] = 1; m->dnum += 12; } /* deltaMax .. f0_old "0", kc fo "1": :83-94 */
  memset(m->nflag, 0, sizeof m->nflag);
  if (m->cohesion) { /* bondFlag, initial_dist, contactPos[3] (newtonflag 0) ; ft, torque/theta n, t (1) ; nonlinear: 14 more trackers (0) */
    m->off_bond = m->dnum; for (int k = 5; k < 14; k++) m->nflag[m->dnum + k] = 1;
    m->dnum += (m->cohesion == C_BOND) ? 14 : 28;
  }
  if (m->tangential) { const int nt = m->tangential == 2 ? 7 : 3; /* hysteretic/nonlinear: shear xyz, shrmag_0, sh_0..2, all "1" (tangential_model_hysteretic_nonlinear.h:79-85) */
    m->off_shear = m->dnum; for (int k = 0; k < nt; k++) m->nflag[m->dnum + k] = 1; m->dnum += nt; }
  if (m->rolling == R_EPSD || m->rolling == R_EPSD2) { m->off_roll = m->dnum; for (int k = 0; k < 3; k++) m->nflag[m->dnum + k] = 1; m->dnum += 3; }
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
    else if (m->cohesion && !strcmp(a[0], "stressBreak")) m->stressBreak = on;
    else if (m->cohesion && !strcmp(a[0], "tensionStress")) m->tension = on;
    else if (m->cohesion && !strcmp(a[0], "compressionStress")) m->compression = on;
    else if (m->cohesion && !strcmp(a[0], "shearStress")) m->shear = on;
    else if (m->cohesion && !strcmp(a[0], "normalTorqueStress")) m->ntorque = on;
    else if (m->cohesion && !strcmp(a[0], "shearTorqueStress")) m->ttorque = on;
    else if (m->cohesion && !strcmp(a[0], "createBondAlways")) m->createAlways = on;
    else if (m->cohesion && !strcmp(a[0], "dampingBond")) m->damping = on;
    else if (m->cohesion == C_BOND && !strcmp(a[0], "dissipationBond")) m->dissipation = on;
    else if (m->cohesion && !strcmp(a[0], "dampingBondSmooth")) m->dampingSmooth = on;
    else if (m->cohesion == C_BOND && !strcmp(a[0], "ratioTensionCompression")) m->ratioTC = on;
    else if (m->cohesion == C_BONDNL && !strcmp(a[0], "ratioTensionCompressionBond")) m->ratioTC = on;
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
  if (w->m.cohesion) return fail(e, "bond models on walls are not supported");
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
int orc_set_extra_force(orc_engine *e, const char *id, int kind, int bit, const double *v, int n)
{
  int k = 0; for (; k < e->nxf; k++) if (!strcmp(e->xf[k].id, id)) break;
  if (n < 0) { if (k == e->nxf) return fail(e, "Could not find fix ID to delete"); for (; k + 1 < e->nxf; k++) e->xf[k] = e->xf[k + 1]; e->nxf--; return 0; }
  if ((kind != 0 && kind != 1) || n != (kind == 0 ? 3 : 1)) return fail(e, "extra force arguments");
  if (k == e->nxf) { if (e->nxf >= 4) return fail(e, "more than 4 fix addforce / viscous"); e->nxf++; snprintf(e->xf[k].id, sizeof e->xf[k].id, "%s", id); }
  e->xf[k].kind = kind; e->xf[k].bit = bit; e->xf[k].v[0] = v[0]; e->xf[k].v[1] = n > 1 ? v[1] : 0.0; e->xf[k].v[2] = n > 2 ? v[2] : 0.0;
  return 0;
}
int orc_set_integrate(orc_engine *e, int bit) { e->integbit = bit; return 0; }

/* mass of a sphere created by fix insert/*: fix_template_sphere.cpp:349-350 (volume_ins = r*r*r*4.*M_PI/3., mass_ins = density_ins*volume_ins),
 * stored by ParticleToInsert::insert particleToInsert.cpp:127 */
static double insert_mass(double r, double rho) { const double vol = r * r * r * 4. * M_PI / 3.; return rho * vol; }

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
    e->rmass[i] = e->ins_mass ? insert_mass(radius[i], density[i]) : 4.0 * M_PI / 3.0 * radius[i] * radius[i] * radius[i] * density[i];
  }
  return 0;
}

/* particles added between two runs (atom_vec_sphere.cpp create_atom via create_atoms / fix insert/*: appended behind the owned
 * atoms, zero velocity unless given, zero force, no contact partners, no wall history); the next orc_setup rebuilds the lists */
int orc_insert_particles(orc_engine *e, long n, const int *tag, const int *type, const int *mask,
                         const double *x, const double *v, const double *omega, const double *radius, const double *density)
{
  if (!e->tag) return orc_upload_particles(e, n, tag, type, mask, x, v, omega, radius, density);
  if (e->nmwalls) return fail(e, "insertion with mesh walls is not restated in the oracle");
  const long n0 = e->n, nn = n0 + n;
#define GROW(p, T, c) do { p = (T *)realloc(p, sizeof(T) * (size_t)(nn ? nn : 1) * (c)); memset(p + (size_t)n0 * (c), 0, sizeof(T) * (size_t)n * (c)); } while (0)
  GROW(e->tag, int, 1); GROW(e->type, int, 1); GROW(e->mask, int, 1); GROW(e->x, double, 3); GROW(e->v, double, 3);
  GROW(e->f, double, 3); GROW(e->omega, double, 3); GROW(e->torque, double, 3); GROW(e->radius, double, 1);
  GROW(e->rmass, double, 1); GROW(e->density, double, 1); GROW(e->xhold, double, 3);
  for (int w = 0; w < e->nwalls; w++) {
    wall_t *W = &e->walls[w]; const int dn = W->m.dnum ? W->m.dnum : 1;
    if (W->hist) GROW(W->hist, double, dn);
    if (W->cand) W->cand = (int *)realloc(W->cand, sizeof(int) * (size_t)(nn ? nn : 1));
  }
  if (e->first) { /* the old half list: no entries for the newcomers */
    e->first = (long *)realloc(e->first, sizeof(long) * (size_t)(nn + 1));
    e->numneigh = (int *)realloc(e->numneigh, sizeof(int) * (size_t)(nn ? nn : 1));
    for (long i = n0; i < nn; i++) { e->first[i] = e->npairs; e->numneigh[i] = 0; }
    e->first[nn] = e->npairs;
  }
#undef GROW
  for (long q = 0; q < n; q++) {
    const long i = n0 + q;
    e->tag[i] = tag[q]; e->type[i] = type[q]; e->mask[i] = mask ? mask[q] : 1;
    for (int d = 0; d < 3; d++) { e->x[3 * i + d] = x[3 * q + d]; e->v[3 * i + d] = v ? v[3 * q + d] : 0.0; e->omega[3 * i + d] = omega ? omega[3 * q + d] : 0.0; }
    e->radius[i] = radius[q]; e->density[i] = density[q];
    e->rmass[i] = e->ins_mass ? insert_mass(radius[q], density[q]) : 4.0 * M_PI / 3.0 * radius[q] * radius[q] * radius[q] * density[q];
  }
  e->n = nn;
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
  double kn, kt, gamman, gammat, Fn, vn, cri, crj, wr1, wr2, wr3, vtr1, vtr2, vtr3, deltaZero;
  double *hist; int *flag;
  double Fi[3], Ti[3], Fj[3], Tj[3];
  double rsq; int has_force_update; long ntimestep; const double *xi;
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

static void normal_hysteretic(const orc_engine *e, const model_t *m, sid_t *s)
{ /* normal_model_hysteretic_nonlinear1.h:149-385 ; variant 2: normal_model_hysteretic_nonlinear2.h (exponential unloading branch,
     square-root damping terms, unloading stiffness from deltaMax) */
  const int V2 = (m->normal == N_HYST2);
  const int it = s->itype, jt = s->jtype;
  const double deltan = s->deltan;
  const double meff = s->meff;
  double kn = e->bp[HP_KEL][it][jt];
  const double Alpha_in = e->bp[HP_ALPHA][it][jt], Cin_in = e->bp[HP_CIN][it][jt], A1_in = e->bp[HP_A1][it][jt], A2_in = e->bp[HP_A2][it][jt];
  const double A3_in = e->bp[HP_A3][it][jt], kcin_in = e->bp[HP_KCIN][it][jt];
  const double k_c = kcin_in * A2_in;
  const double kc = e->bp[HP_KN2KC][it][jt] * kn;
  const double f_0 = e->bp[HP_FADH][it][jt];
  const double crl = e->corLog[it][jt];
  double gamman, gammat;
  if (s->flag) *s->flag |= CONTACT_NORMAL;
  double *history = &s->hist[m->off_norm];
  double deltaMax;
  if (deltan > history[0]) { history[0] = deltan; deltaMax = deltan; } else deltaMax = history[0];
  double deltaZero = history[1];
  double k1 = history[2];
  double deltaZero_old = history[3];
  double k1_old = history[4];
  const double delta_old = history[5];
  double deltaMin = history[6];
  double betan = history[7];
  double f0 = history[8];
  double f0_old = history[9];
  double k2, fHys;
  k2 = A3_in * k1;
  int tag_status; /* loading 1, unloading 2 */
  if (deltan >= delta_old) {
    if (delta_old == 0) tag_status = 1;
    else tag_status = (s->vn > 0) ? 2 : 1;
  } else tag_status = 2;
  const double sq54 = sqrt(5 / 4); /* integer division in the reference: sqrt(1) */
  const double dexp = V2 ? 0.5 : 0.25;
  const double cdamp = 1. + (M_PI / crl) * (M_PI / crl);
  if (tag_status == 1) {
    if (deltaZero == 0) k1 = A2_in; else k1 = history[2];
    deltaMin = history[6];
    f0_old = f0;
    if (deltan <= deltaZero) {
      if (!V2) { if (deltan <= deltaMin) fHys = -k_c * deltan; else fHys = betan * (deltan - deltaZero); }
      else fHys = Cin_in * k2 * (exp(betan * (deltan - deltaZero)) - 1);
    } else fHys = Alpha_in * k1 * pow(deltan - deltaZero, 2) + f0;
    gamman = sq54 * sqrt(4. * meff * Alpha_in * k1 / cdamp) * (pow(deltan, dexp) + pow(deltaZero, dexp));
    gammat = gamman;
    history[1] = deltaZero; history[2] = k1; history[3] = deltaZero; history[4] = k1; history[5] = deltan;
    history[6] = deltaMin; history[7] = betan; history[8] = f0; history[9] = f0_old;
  } else if (!V2) {
    k2 = A3_in * k1_old;
    deltaZero = (1 - k1_old / k2) * deltaMax;
    const double beta = Alpha_in * k1_old * pow(deltaMax - deltaZero_old, 2) / k2 / (deltaMax - deltaZero);
    deltaMin = beta * (k2 - k1_old) / (beta * k2 + k_c) * deltaMax;
    k1 = deltaMax * A1_in + A2_in;
    if (deltan >= deltaMin) {
      betan = beta * k2;
      if (deltan >= deltaZero) { fHys = beta * k2 * (deltan - deltaZero) + (deltan - deltaZero) * f0_old / (deltaMax - deltaZero); f0 = fHys - Alpha_in * k1 * pow(deltan - deltaZero, 2); }
      else { fHys = beta * k2 * (deltan - deltaZero); f0 = 0; }
    } else { fHys = -k_c * deltan; f0 = 0; }
    gamman = sq54 * sqrt(4. * meff * Alpha_in * k1_old / cdamp) * pow(deltan, 0.25);
    gammat = gamman;
    history[1] = deltaZero; history[2] = k1; history[3] = deltaZero_old; history[4] = k1_old; history[5] = deltan;
    history[6] = deltaMin; history[7] = betan; history[8] = f0;
  } else {
    const double fcc = 1.0;
    k2 = A3_in * (A1_in * deltaMax + A2_in);
    deltaZero = fcc * (1 - k1_old / k2) * deltaMax;
    deltaZero_old = history[3];
    betan = log(Alpha_in * k1_old / Cin_in / k2 * pow(deltaMax - deltaZero_old, 2) + 1) / (deltaMax - fcc * (1 - k1_old / k2) * deltaMax);
    deltaMin = betan * (k2 - k1_old) / (betan * k2 + k_c) * deltaMax;
    k1 = deltaMax * A1_in + A2_in;
    if (deltan >= deltaZero) { fHys = Cin_in * k2 * (exp(betan * (deltan - deltaZero)) - 1) + (deltan - deltaZero) * f0_old / (deltaMax - deltaZero); f0 = fHys - Alpha_in * k1 * pow(deltan - deltaZero, 2); }
    else { fHys = Cin_in * k2 * (exp(betan * (deltan - deltaZero)) - 1); f0 = 0; }
    gamman = 1 * 0.001 * sq54 * sqrt(4. * meff * Alpha_in * k1_old / cdamp) * pow(deltan, -0.25);
    gammat = gamman;
    history[1] = deltaZero; history[2] = k1; history[3] = deltaZero_old; history[4] = k1_old; history[5] = deltan;
    history[6] = deltaMin; history[7] = betan; history[8] = f0;
  }
  kn = k1;
  double kt = kn;
  kn /= e->nktv2p; kt /= e->nktv2p;
  const double Fn_damping = -gamman * s->vn;
  double Fn = fHys + Fn_damping + f_0;
  if (m->limitForce && (Fn < 0.0) && kc == 0 && f_0 == 0.0) Fn = 0.0;
  /* (the model registers tangential_damping but never applies it: gammat is used as computed, :186-191 are commented out) */
  s->Fn = Fn; s->kn = kn; s->kt = kt; s->gamman = gamman; s->gammat = gammat; s->deltaZero = deltaZero;
  history[10] = kc; history[11] = f_0;
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
  double shrmag = sqrt(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2]);
  const double kt = s->kt;
  const double xmu = e->mu[s->itype][s->jtype];
  if (m->tangential == 2) { /* tangential_model_hysteretic_nonlinear.h:186-206: inside the plastic range of the normal law the spring restarts */
    double shrmag_0 = shear[3], sh_0 = shear[4], sh_1 = shear[5], sh_2 = shear[6];
    if (s->deltan <= s->deltaZero) { shrmag_0 = shrmag; shear[3] = shrmag_0; sh_0 = shear[0]; sh_1 = shear[1]; sh_2 = shear[2]; }
    shear[0] -= sh_0; shear[1] -= sh_1; shear[2] -= sh_2;
    shrmag -= shrmag_0;
  }
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
      const double Fn = m->cdtnl2 ? s->Fn : s->deltan * s->kn; /* rolling_model_cdtnonlinear2.h:127 */
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
      const double sc = m->cdtnl2 ? rmu * s->Fn * reff / mag : rmu * s->kn * s->deltan * reff / mag; /* rolling_model_cdtnonlinear2.h:157 */
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


/* ---------------------------------------------------------------- cohesion models bond and bond/nonlinear (sphere-sphere) */
static void vproject(const double *v, const double *on, double *res)
{ /* vector_liggghts.h:437-442 vectorProject3D: normalises `on` first (zero vector -> zero) */
  double n[3] = {on[0], on[1], on[2]};
  const double norm = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]); const double inv = (norm == 0.) ? 0. : 1. / norm;
  n[0] *= inv; n[1] *= inv; n[2] *= inv;
  const double d = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
  res[0] = n[0] * d; res[1] = n[1] * d; res[2] = n[2] * d;
}
static int isgn(double v) { return (0. < v) - (v < 0.); } /* math_extra_liggghts.h:120-123 */
static double damp_mult(const model_t *m, double vel, double minvel, double dv) { return m->dampingSmooth ? fmin(1.0, fmax(-1.0, vel / fmax(minvel, dv))) : (double)isgn(vel); }

/* hysteretic torque law of one component, cohesion_model_bond_nonlinear.h:669-872; tmax/tmin = the running extreme angles */
static double nl_torque_comp(double theta, double *tmax, double *tmin, double k, double ku, double kc, double JI)
{
  if (theta > 0.0) *tmin = 0.0; else if (theta < 0.0) *tmax = 0.0;
  const double c1 = (ku - k) / (ku + kc) * *tmax, c2 = (ku - k) / (ku + kc) * *tmin;
  double tq;
  if ((theta >= *tmax) || (theta <= *tmin)) tq = -k * JI * theta;
  else if (theta > c1) tq = -ku * JI * theta + (ku - k) * JI * *tmax;
  else if (theta >= c2) tq = kc * JI * theta;
  else tq = -ku * JI * theta + (ku - k) * JI * *tmin;
  if (theta > *tmax) *tmax = theta;
  if (theta < *tmin) *tmin = theta;
  return tq;
}

/* CohesionModel<COHESION_BOND>::surfacesClose cohesion_model_bond.h:491-950 and
 * CohesionModel<COHESION_BOND_NONLINEAR>::surfacesClose cohesion_model_bond_nonlinear.h:394-940, sphere-sphere branch.
 * Called for touching pairs (via surfacesIntersect) and for pairs inside the contact-distance band. */
static void cohesion_bond(const orc_engine *e, const model_t *m, sid_t *s)
{
  const int it = s->itype, jt = s->jtype, NL = (m->cohesion == C_BONDNL);
  const int update_history = s->shearupdate; /* computeflag is 1 on this path */
  const double lambda = e->bp[BP_LAMBDA][it][jt];
  if (lambda < 1.e-15) return;
  double *H = &s->hist[m->off_bond];
  const double r_create = sqrt(s->rsq);
  double r = r_create;
  if (update_history) {
    if (H[0] < 1.e-15) {
      const int create = (m->createAlways || (s->ntimestep == (int)e->tsCreateBond)) && (r_create < e->bp[BP_CREATEDIST][it][jt]);
      if (!create) return;
      /* createBond :1015-1034 / createBondnonlinear :958-975 */
      if (s->flag) *s->flag |= CONTACT_COHESION;
      H[0] = 1.0; H[1] = r;
      for (int d = 0; d < 3; d++) H[2 + d] = s->xi[d] - s->delta[d];
      for (int d = 5; d < 14; d++) H[d] = 0.0;
      ((orc_engine *)e)->bond_created++; /* cohesion_model_bond.h:1032 */
    }
  } else if (H[0] < 1.e-15) return;
  double force_tang[3] = {H[5], H[6], H[7]}, tn[3] = {H[8], H[9], H[10]}, tt[3] = {H[11], H[12], H[13]}; /* linear: torques ; nonlinear: angles */
  const double radi = s->radi, radj = s->radj;
  const double *delta = s->delta;
  if (!m->stressBreak && r > e->bp[BP_MAXDIST][it][jt] && update_history) { /* breakBond :1036-1070 */
    if (s->flag) *s->flag &= ~CONTACT_COHESION;
    ((orc_engine *)e)->bond_broken++; /* :1066 */
    H[0] = 0.; H[1] = 0.; return;
  }
  if (s->flag) *s->flag |= CONTACT_COHESION;
  const double rinv = 1. / r;
  const double en[3] = {delta[0] * rinv, delta[1] * rinv, delta[2] * rinv};
  const double *vi = s->vi, *vj = s->vj, *omegai = s->wi, *omegaj = s->wj;
  const double rb = lambda * (radi < radj ? radi : radj);
  const double A = M_PI * rb * rb, J = 0.5 * A * rb * rb, I = 0.5 * J, dt = e->dt;
  const double radsuminv = 1. / (radi + radj);
  const double cri = r * radi * radsuminv, crj = r * radj * radsuminv;
  double vr[3], vn[3], vt[3], wr[3], tmp1[3], tmp2[3], vtr[3], wn[3], wt[3];
  for (int d = 0; d < 3; d++) vr[d] = vi[d] - vj[d];
  vproject(vr, en, vn);
  for (int d = 0; d < 3; d++) vt[d] = vr[d] - vn[d];
  for (int d = 0; d < 3; d++) { tmp1[d] = omegai[d] * (radi * radsuminv); tmp2[d] = omegaj[d] * (radj * radsuminv); wr[d] = tmp1[d] + tmp2[d]; }
  tmp1[0] = delta[1] * wr[2] - delta[2] * wr[1]; tmp1[1] = delta[2] * wr[0] - delta[0] * wr[2]; tmp1[2] = delta[0] * wr[1] - delta[1] * wr[0];
  for (int d = 0; d < 3; d++) vtr[d] = vt[d] + tmp1[d];
  for (int d = 0; d < 3; d++) wr[d] = omegai[d] - omegaj[d];
  vproject(wr, en, wn);
  for (int d = 0; d < 3; d++) wt[d] = wr[d] - wn[d];
  double nforce[3] = {0., 0., 0.}, nforce_d[3] = {0., 0., 0.}, tforce_d[3] = {0., 0., 0.}, ntorque_d[3] = {0., 0., 0.}, ttorque_d[3] = {0., 0., 0.};
  double torque_normal[3] = {0., 0., 0.}, torque_tang[3] = {0., 0., 0.};
  const double displacement = H[1] - r;
  const double minvel = 1e-5 * fmin(radi, radj) / dt;
  const double dfn = e->bp[BP_DFN][it][jt], dft = e->bp[BP_DFT][it][jt], dtn = e->bp[BP_DTN][it][jt], dtt = e->bp[BP_DTT][it][jt];
  double dmax = 0., dmin = 0.;
  if (NL) { /* running extreme displacements :566-580 (updated even when shearupdate == 0) */
    if (displacement > H[14]) { H[14] = displacement; dmax = displacement; } else dmax = H[14];
    if (displacement < H[27]) { H[27] = displacement; dmin = displacement; } else dmin = H[27];
    if (displacement < 0.0) H[14] = 0.0;
    if (displacement > 0.0) H[27] = 0.0;
  }
  if (m->tension || m->compression) {
    if (!NL && m->dissipation && update_history) { /* relax the normal spring, cohesion_model_bond.h:677-687 (the force below still uses the displacement formed before) */
      const double dissipate = fmin(dt * (1. / e->bp[DP_FN][it][jt]), 1.0);
      H[1] += (r - H[1]) * dissipate;
    }
    if ((m->tension && displacement < -1.e-15) || (m->compression && displacement > 1.e-15)) {
      double frcmag;
      if (!NL) frcmag = e->bp[BP_KN][it][jt] * A * displacement;
      else {
        const double k1 = e->bp[BP_K_FN1][it][jt], ku1 = e->bp[BP_KU_FN1][it][jt], kc1 = e->bp[BP_KC_FN1][it][jt];
        const double k2 = e->bp[BP_K_FN2][it][jt], ku2 = e->bp[BP_KU_FN2][it][jt], kc2 = e->bp[BP_KC_FN2][it][jt];
        const double c1 = pow((ku1 - k1) / (ku1 + kc1), 2.0) * dmax, c2 = pow((ku2 - k2) / (ku2 + kc2), 1.0) * dmin;
        if (displacement < dmin) frcmag = k2 * A * displacement;
        else if (displacement < c2) frcmag = ku2 * A * displacement + (k2 - ku2) * A * dmin;
        else if (displacement < 0.0) frcmag = -kc2 * A * displacement;
        else if (displacement < c1) frcmag = -kc1 * pow(fabs(displacement), 0.5);
        else if (displacement < dmax) frcmag = ku1 * pow(fabs(displacement), 0.5) + (k1 - ku1) * pow(fabs(dmax), 0.5);
        else frcmag = k1 * pow(fabs(displacement), 0.5);
      }
      for (int d = 0; d < 3; d++) nforce[d] = en[d] * frcmag;
      if (m->damping) for (int d = 0; d < 3; d++) nforce_d[d] = nforce[d] - dfn * fabs(nforce[d]) * damp_mult(m, vn[d], minvel, 0.01 * nforce[d] * dt);
      else for (int d = 0; d < 3; d++) nforce_d[d] = nforce[d];
    }
  }
  if (m->shear) {
    const double ktA = (NL ? e->bp[BP_K_FT][it][jt] : e->bp[BP_KT][it][jt]);
    double dtforce[3]; for (int d = 0; d < 3; d++) dtforce[d] = vtr[d] * (-ktA * A * dt);
    vproject(force_tang, en, tmp1);
    for (int d = 0; d < 3; d++) force_tang[d] = force_tang[d] - tmp1[d];
    if (!NL && m->dissipation) { const double k = 1.0 - fmin(dt * (1. / e->bp[DP_FT][it][jt]), 1.0); for (int d = 0; d < 3; d++) force_tang[d] = force_tang[d] * k; } /* :731-732 */
    for (int d = 0; d < 3; d++) force_tang[d] = force_tang[d] + dtforce[d];
    if (m->damping) for (int d = 0; d < 3; d++) tforce_d[d] = force_tang[d] - dft * fabs(force_tang[d]) * damp_mult(m, vtr[d], minvel, 0.01 * force_tang[d] * dt);
    else for (int d = 0; d < 3; d++) tforce_d[d] = force_tang[d];
  }
  if (!NL) {
    const double kn_pb = e->bp[BP_KN][it][jt], kt_pb = e->bp[BP_KT][it][jt];
    for (int d = 0; d < 3; d++) { torque_normal[d] = tn[d]; torque_tang[d] = tt[d]; }
    if (m->ntorque) {
      double dnt[3]; for (int d = 0; d < 3; d++) dnt[d] = wn[d] * (-kt_pb * J * dt);
      vproject(torque_normal, en, torque_normal);
      if (m->dissipation) { const double k = 1.0 - fmin(dt * (1. / e->bp[DP_TN][it][jt]), 1.0); for (int d = 0; d < 3; d++) torque_normal[d] = torque_normal[d] * k; } /* :763-764 */
      for (int d = 0; d < 3; d++) torque_normal[d] = torque_normal[d] + dnt[d];
      if (m->damping) for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d] - dtn * fabs(torque_normal[d]) * isgn(wn[d]);
      else for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d];
    }
    if (m->ttorque) {
      const double wtsq = wt[0] * wt[0] + wt[1] * wt[1] + wt[2] * wt[2];
      if (wtsq > 0) {
        double dtt3[3]; for (int d = 0; d < 3; d++) dtt3[d] = wt[d] * (-kn_pb * I * dt);
        vproject(torque_tang, wt, torque_tang);
        if (m->dissipation) { const double k = 1.0 - fmin(dt * (1. / e->bp[DP_TT][it][jt]), 1.0); for (int d = 0; d < 3; d++) torque_tang[d] = torque_tang[d] * k; } /* :798-799 */
        for (int d = 0; d < 3; d++) torque_tang[d] = torque_tang[d] + dtt3[d];
        if (m->damping) for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d] - dtt * fabs(torque_tang[d]) * isgn(wt[d]);
        else for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d];
      }
    }
  } else {
    if (m->ntorque) { /* :658-757 */
      const double k = e->bp[BP_K_TN][it][jt], ku = e->bp[BP_KU_TN][it][jt], kc = e->bp[BP_KC_TN][it][jt];
      for (int d = 0; d < 3; d++) tn[d] = tn[d] + wn[d] * dt;
      for (int d = 0; d < 3; d++) torque_normal[d] = tn[d] * (-k * J);
      vproject(torque_normal, wn, torque_normal);
      for (int d = 0; d < 3; d++) torque_normal[d] = nl_torque_comp(tn[d], &H[15 + d], &H[21 + d], k, ku, kc, J);
      if (m->damping) for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d] - dtn * fabs(torque_normal[d]) * isgn(wn[d]);
      else for (int d = 0; d < 3; d++) ntorque_d[d] = torque_normal[d];
    }
    if (m->ttorque) { /* :760-872 */
      const double k = e->bp[BP_K_TT][it][jt], ku = e->bp[BP_KU_TT][it][jt], kc = e->bp[BP_KC_TT][it][jt];
      for (int d = 0; d < 3; d++) tt[d] = tt[d] + wt[d] * dt;
      for (int d = 0; d < 3; d++) torque_tang[d] = nl_torque_comp(tt[d], &H[18 + d], &H[24 + d], k, ku, kc, I);
      if (m->damping) for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d] - dtt * fabs(torque_tang[d]) * isgn(wt[d]);
      else for (int d = 0; d < 3; d++) ttorque_d[d] = torque_tang[d];
    }
  }
  if (m->stressBreak) { /* :816-848 / :875-890 (un-damped forces and torques) */
    const double nfm = sqrt(nforce[0] * nforce[0] + nforce[1] * nforce[1] + nforce[2] * nforce[2]), tfm = sqrt(force_tang[0] * force_tang[0] + force_tang[1] * force_tang[1] + force_tang[2] * force_tang[2]),
      ntm = sqrt(torque_normal[0] * torque_normal[0] + torque_normal[1] * torque_normal[1] + torque_normal[2] * torque_normal[2]), ttm = sqrt(torque_tang[0] * torque_tang[0] + torque_tang[1] * torque_tang[1] + torque_tang[2] * torque_tang[2]);
    double maxSigma = e->bp[BP_MAXSIGMA][it][jt];
    if (m->ratioTC && (NL ? displacement < -1.e-15 : displacement < 1e-16)) maxSigma *= e->bp[BP_RATIOTC][it][jt];
    const int nstress = maxSigma < (nfm / A + ttm * rb / I);
    const int tstress = e->bp[BP_MAXTAU][it][jt] < (tfm / A + ntm * rb / J);
    if ((nstress || tstress) && update_history) { if (s->flag) *s->flag &= ~CONTACT_COHESION; ((orc_engine *)e)->bond_broken++; H[0] = 0.; H[1] = 0.; return; }
  }
  double tor[3]; tor[0] = tforce_d[1] * en[2] - tforce_d[2] * en[1]; tor[1] = tforce_d[2] * en[0] - tforce_d[0] * en[2]; tor[2] = tforce_d[0] * en[1] - tforce_d[1] * en[0];
  s->has_force_update = 1;
  double F[3], Ti[3], Tj[3];
  for (int d = 0; d < 3; d++) { F[d] = nforce_d[d] + tforce_d[d]; Ti[d] = cri * tor[d] + ntorque_d[d] + ttorque_d[d]; Tj[d] = crj * tor[d] - ntorque_d[d] - ttorque_d[d]; }
  if (update_history) for (int d = 0; d < 3; d++) { H[5 + d] = force_tang[d]; H[8 + d] = NL ? tn[d] : torque_normal[d]; H[11 + d] = NL ? tt[d] : torque_tang[d]; }
  if (!NL) for (int d = 0; d < 3; d++) { s->Fi[d] += F[d]; s->Ti[d] += Ti[d]; s->Fj[d] -= F[d]; s->Tj[d] += Tj[d]; }
  else for (int d = 0; d < 3; d++) { s->Fi[d] = F[d]; s->Ti[d] = Ti[d]; s->Fj[d] = -F[d]; s->Tj[d] = Tj[d]; } /* the nonlinear bond OVERWRITES :917-937 */
}

/* ContactModel::surfacesIntersect, contact_models.h:228-238 */
static void chain_intersect(const orc_engine *e, const model_t *m, sid_t *s)
{
  surface_default(s);
  if (m->normal == N_HERTZ) normal_hertz(e, m, s); else if (m->normal == N_HOOKE) normal_hooke(e, m, s); else normal_hysteretic(e, m, s);
  if (m->cohesion) cohesion_bond(e, m, s);
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


/* ================================================================ triangle mesh walls
 * fix mesh/surface + fix wall/gran ... mesh: geometry, topology, candidate lists, contact history. */
static void v3sub(const double *a, const double *b, double *r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static double v3dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3mag(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static void v3cross(const double *a, const double *b, double *r) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; }
static void v3sdiv(double *v, double s) { const double sinv = 1. / s; v[0] = sinv * v[0]; v[1] = sinv * v[1]; v[2] = sinv * v[2]; } /* vector_liggghts.h:288-294 */
static int comp_double(double a, double b, double prec)
{ /* math_extra_liggghts.h:560-571 */
  if (a == b) return 1;
  if (b == 0) return a < prec && a > -prec;
  const double x = (a - b);
  return x < prec && x > -prec;
}
static int nodes_equal(const mesh_t *M, const double *a, const double *b)
{ for (int d = 0; d < 3; d++) if (!comp_double(a[d], b[d], M->precision)) return 0; return 1; } /* multi_node_mesh_I.h:246-261 */

static void tri_surf_properties(mesh_t *M, int n);
static void tri_properties(mesh_t *M, int n)
{ /* multi_node_mesh_I.h:153-172 (center, rBound) ; surface_mesh_I.h:302-470 */
  double avg[3] = {0., 0., 0.};
  for (int i = 0; i < 3; i++) for (int d = 0; d < 3; d++) avg[d] = M->node[n][i][d] + avg[d];
  v3sdiv(avg, 3.0);
  for (int d = 0; d < 3; d++) M->center[n][d] = avg[d];
  double rb = 0.;
  for (int i = 0; i < 3; i++) { double vec[3]; v3sub(M->center[n], M->node[n][i], vec); const double m = v3mag(vec); if (m > rb) rb = m; }
  M->rbound[n] = rb;
  tri_surf_properties(M, n);
}
static void tri_surf_properties(mesh_t *M, int n)
{ /* SurfaceMesh::recalcLocalSurfProperties surface_mesh_I.h:187-222: edge vectors/lengths, surface and edge normals, obtuse
   * index recomputed from the nodes -- at first setup, at every later setup and, for moving meshes, at every neighbour
   * rebuild (FixMesh::setup_pre_force / pre_force fix_mesh.cpp:491-575 -> pbcExchangeBorders -> refreshOwned) */
  for (int i = 0; i < 3; i++) { /* calcEdgeVecLen */
    v3sub(M->node[n][(i + 1) % 3], M->node[n][i], M->edgeVec[n][i]);
    M->edgeLen[n][i] = v3mag(M->edgeVec[n][i]);
    v3sdiv(M->edgeVec[n][i], M->edgeLen[n][i]);
  }
  double *sn = M->surfNorm[n]; /* calcSurfaceNorm (non-degenerate branch) */
  v3cross(M->edgeVec[n][0], M->edgeVec[n][1], sn);
  v3sdiv(sn, v3mag(sn));
  for (int i = 0; i < 3; i++) { /* calcEdgeNormals */
    v3cross(M->edgeVec[n][i], sn, M->edgeNorm[n][i]);
    v3sdiv(M->edgeNorm[n][i], v3mag(M->edgeNorm[n][i]));
  }
  /* calcObtuseAngleIndex is called for iNode = 0,1,2 and each call overwrites the value (surface_mesh_I.h:331-339,
   * 455-467): what survives is the verdict of node 2 */
  M->obtuse[n] = -1;
  for (int i = 0; i < 3; i++) { const double dot = v3dot(M->edgeVec[n][i], M->edgeVec[n][(i - 1 + 3) % 3]); M->obtuse[n] = dot > 0. ? i : -1; }
}

static int share_edge(const mesh_t *M, int i, int j, int *iEdge, int *jEdge)
{ /* multi_node_mesh_I.h:281-321 share2Nodes + surface_mesh_I.h:1040-1066 shareEdge */
  double dist[3]; v3sub(M->center[i], M->center[j], dist);
  const double radsum = M->rbound[i] + M->rbound[j];
  if (v3dot(dist, dist) > radsum * radsum) return 0;
  int nShared = 0, i1 = -1, j1 = -1, i2 = -1, j2 = -1;
  for (int a = 0; a < 3 && i2 < 0; a++) for (int b = 0; b < 3; b++) if (nodes_equal(M, M->node[i][a], M->node[j][b])) {
    if (nShared == 0) { i1 = a; j1 = b; } else { i2 = a; j2 = b; break; }
    nShared++;
  }
  if (i2 < 0) return 0;
  *iEdge = (2 == i1 + i2) ? 2 : (i1 < i2 ? i1 : i2);
  *jEdge = (2 == j1 + j2) ? 2 : (j1 < j2 ? j1 : j2);
  return 1;
}
static int n_shared_nodes(const mesh_t *M, int i, int j)
{ int n = 0; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) if (nodes_equal(M, M->node[i][a], M->node[j][b])) n++; return n; }

static int coplanar_neighs_overlap(const mesh_t *M, int i, int iEdge, int j, int jEdge)
{ /* surface_mesh_I.h:965-1003 */
  double vecI[3], vecJ[3];
  const double *pRef = M->node[i][iEdge], *edgeN = M->edgeNorm[i][iEdge];
  v3sub(M->node[i][(iEdge + 2) % 3], pRef, vecI); v3sub(M->node[j][(jEdge + 2) % 3], pRef, vecJ);
  return v3dot(vecI, edgeN) * v3dot(vecJ, edgeN) > 0.;
}
static void handle_shared_edge(mesh_t *M, int i, int iEdge, int j, int jEdge, int coplanar)
{ /* surface_mesh_I.h:1071-1144 (ids == indices; the *Active_ override flags are all false) */
  if (M->nNeighs[i] < NUM_NEIGH_MAX) M->neighFaces[i][M->nNeighs[i]] = j;
  if (M->nNeighs[j] < NUM_NEIGH_MAX) M->neighFaces[j][M->nNeighs[j]] = i;
  M->nNeighs[i]++; M->nNeighs[j]++;
  if (!coplanar || coplanar_neighs_overlap(M, i, iEdge, j, jEdge)) {
    if (i < j) { M->edgeActive[i][iEdge] = 0; M->edgeActive[j][jEdge] = 1; }
    else { M->edgeActive[i][iEdge] = 1; M->edgeActive[j][jEdge] = 0; }
  } else { M->edgeActive[i][iEdge] = 0; M->edgeActive[j][jEdge] = 0; }
}
typedef struct { int nVisited, nHasNode, anyActive; int *visited, *hasNode; double (*edgeList)[3], (*endPoint)[3]; } corner_t;
static void check_node_recursive(const mesh_t *M, int iSrf, const double *node, corner_t *C)
{ /* surface_mesh_I.h:1192-1236 */
  for (int k = 0; k < C->nVisited; k++) if (C->visited[k] == iSrf) return;
  C->visited[C->nVisited++] = iSrf;
  int iNode = -1;
  for (int a = 0; a < 3 && iNode < 0; a++) if (nodes_equal(M, M->node[iSrf][a], node)) iNode = a; /* containsNode */
  if (iNode < 0) return;
  int ne = 2 * C->nHasNode;
  C->hasNode[C->nHasNode++] = iSrf;
  const int im = (iNode - 1 + 3) % 3;
  memcpy(C->edgeList[ne], M->edgeVec[iSrf][iNode], 24); memcpy(C->endPoint[ne++], M->node[iSrf][(iNode + 1) % 3], 24);
  memcpy(C->edgeList[ne], M->edgeVec[iSrf][im], 24); memcpy(C->endPoint[ne++], M->node[iSrf][im], 24);
  if (M->edgeActive[iSrf][iNode]) C->anyActive = 1; else if (M->edgeActive[iSrf][im]) C->anyActive = 1;
  const int nn = M->nNeighs[iSrf] < NUM_NEIGH_MAX ? M->nNeighs[iSrf] : NUM_NEIGH_MAX;
  for (int k = 0; k < nn; k++) { const int idN = M->neighFaces[iSrf][k]; if (idN < 0) return; check_node_recursive(M, idN, node, C); }
}
static int in_subdomain(const orc_engine *e, const double *pos)
{ /* domain_I.h:65-87, one rank: sub-box == box, padded by SMALL_DMBRDR = 1e-8 */
  for (int d = 0; d < 3; d++) if (!(pos[d] >= e->lo[d] - 1.0e-8 && pos[d] < e->hi[d] + 1.0e-8)) return 0;
  return 1;
}
static void mesh_topology(const orc_engine *e, mesh_t *M)
{ /* SurfaceMesh::buildNeighbours surface_mesh_I.h:474-582 */
  const int n = M->ntri;
  for (int i = 0; i < n; i++) {
    M->nNeighs[i] = 0;
    for (int k = 0; k < NUM_NEIGH_MAX; k++) M->neighFaces[i][k] = -1;
    for (int k = 0; k < 3; k++) { M->edgeActive[i][k] = 1; M->cornerActive[i][k] = 1; }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) { /* elements inserted before i (:519-545) */
    int iEdge = 0, jEdge = 0;
    if (share_edge(M, i, j, &iEdge, &jEdge)) {
      const double dot = v3dot(M->surfNorm[i], M->surfNorm[j]);
      handle_shared_edge(M, i, iEdge, j, jEdge, fabs(dot) >= M->curvature); /* areCoplanar :874-890 */
    }
  }
  corner_t C; C.visited = (int *)malloc(sizeof(int) * (n + 1)); C.hasNode = (int *)malloc(sizeof(int) * (n + 1));
  C.edgeList = (double (*)[3])malloc(sizeof(double) * 3 * 2 * (n + 1)); C.endPoint = (double (*)[3])malloc(sizeof(double) * 3 * 2 * (n + 1));
  for (int i = 0; i < n; i++) for (int iNode = 0; iNode < 3; iNode++) { /* handleCorner :1148-1188 */
    C.nVisited = C.nHasNode = C.anyActive = 0;
    check_node_recursive(M, i, M->node[i][iNode], &C);
    if (!in_subdomain(e, M->node[i][iNode])) continue;
    int maxId = -1; for (int k = 0; k < C.nHasNode; k++) if (C.hasNode[k] > maxId) maxId = C.hasNode[k];
    int colinear = 0; const int nE = 2 * C.nHasNode;
    for (int a = 0; a < nE; a++) for (int b = a + 1; b < nE; b++)
      if (fabs(v3dot(C.edgeList[a], C.edgeList[b])) > M->curvature && !nodes_equal(M, C.endPoint[a], C.endPoint[b])) colinear = 1; /* edgeVecsColinear :1007-1014 */
    if (colinear || !C.anyActive) M->cornerActive[i][iNode] = 0;
    else if (i == maxId) M->cornerActive[i][iNode] = 1;
    else M->cornerActive[i][iNode] = 0;
  }
  free(C.visited); free(C.hasNode); free(C.edgeList); free(C.endPoint);
}
static int coplanar_node_neighs(const mesh_t *M, int a, int b)
{ /* surface_mesh_I.h:921-958 areCoplanarNodeNeighs */
  int areNeighs = 0;
  const int nn = M->nNeighs[a] < NUM_NEIGH_MAX ? M->nNeighs[a] : NUM_NEIGH_MAX;
  for (int k = 0; k < nn; k++) if (M->neighFaces[a][k] == b) areNeighs = 1;
  if (!areNeighs && n_shared_nodes(M, a, b) == 0) return 0;
  return fabs(v3dot(M->surfNorm[a], M->surfNorm[b])) > M->curvature;
}

int orc_add_mesh(orc_engine *e, const char *id, int atom_type, const double *nodes, long ntri, int argc, const char *const *argv)
{ /* fix mesh/surface file F type T : fix_mesh.cpp:90-260, fix_mesh_surface.cpp:90-330 (nodes are the file's vertices) */
  if (e->nmeshes == MAXMESH) return fail(e, "too many meshes");
  mesh_t *M = &e->meshes[e->nmeshes]; memset(M, 0, sizeof *M);
  snprintf(M->id, sizeof M->id, "%s", id); M->atom_type = atom_type; M->ntri = (int)ntri; M->wall = -1;
  M->curvature = 1. - EPSILON_CURVATURE; M->precision = EPSILON_PRECISION;
  for (int k = 0; k + 1 < argc; k += 2) {
    if (!strcmp(argv[k], "curvature")) M->curvature = cos(atof(argv[k + 1]) * M_PI / 180.); /* fix_mesh.cpp: curvature given in degrees */
    else if (!strcmp(argv[k], "precision")) M->precision = atof(argv[k + 1]);
    else if (!strcmp(argv[k], "stress")) M->stress = !strcmp(argv[k + 1], "on"); /* mesh_module_stress.cpp:120-140 */
    else if (!strcmp(argv[k], "reference_point") && k + 3 < argc) { for (int d = 0; d < 3; d++) M->p_ref[d] = atof(argv[k + 1 + d]); k += 2; }
    else return fail(e, "mesh option not supported by the oracle");
  }
  const size_t T = (size_t)(ntri ? ntri : 1);
  M->node = calloc(T, sizeof *M->node); M->center = calloc(T, sizeof *M->center); M->rbound = calloc(T, sizeof(double));
  M->edgeVec = calloc(T, sizeof *M->edgeVec); M->edgeLen = calloc(T, sizeof *M->edgeLen); M->surfNorm = calloc(T, sizeof *M->surfNorm);
  M->edgeNorm = calloc(T, sizeof *M->edgeNorm); M->obtuse = calloc(T, sizeof(int)); M->nNeighs = calloc(T, sizeof(int));
  M->neighFaces = calloc(T, sizeof *M->neighFaces); M->edgeActive = calloc(T, sizeof *M->edgeActive); M->cornerActive = calloc(T, sizeof *M->cornerActive);
  M->vnode = calloc(T, sizeof *M->vnode); M->nodesLastRe = calloc(T, sizeof *M->nodesLastRe);
  M->contacts = calloc(T, sizeof(int *)); M->ncontacts = calloc(T, sizeof(int)); M->capcontacts = calloc(T, sizeof(int));
  memcpy(M->node, nodes, sizeof(double) * 9 * ntri);
  for (int n = 0; n < ntri; n++) tri_properties(M, n);
  mesh_topology(e, M);
  e->nmeshes++; return 0;
}
int orc_move_mesh(orc_engine *e, const char *mesh_id, int argc, const char *const *argv)
{ /* fix move/mesh mesh ID linear vx vy vz : fix_move_mesh.cpp, mesh_mover_linear.cpp:94-112
   * fix move/mesh mesh ID rotate origin x y z axis x y z period T : mesh_mover_rotation.cpp:58-82 */
  for (int m = 0; m < e->nmeshes; m++) if (!strcmp(e->meshes[m].id, mesh_id)) {
    mesh_t *M = &e->meshes[m];
    if (M->moving) return fail(e, "one fix move/mesh per mesh is supported");
    if (argc == 4 && !strcmp(argv[0], "linear")) {
      for (int d = 0; d < 3; d++) M->vel[d] = atof(argv[1 + d]);
      M->moving = 1;
    } else if (argc >= 11 && !strcmp(argv[0], "rotate")) {
      if (strcmp(argv[1], "origin")) return fail(e, "Expected keyword 'origin'");
      if (strcmp(argv[5], "axis")) return fail(e, "Expected keyword 'axis'");
      if (strcmp(argv[9], "period")) return fail(e, "Expected keyword 'period'");
      for (int d = 0; d < 3; d++) { M->rot_origin[d] = atof(argv[2 + d]); M->rot_axis[d] = atof(argv[6 + d]); }
      { /* vectorNormalize3D vector_liggghts.h:63-70 */
        double *v = M->rot_axis; const double norm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        const double invnorm = (norm == 0.) ? 0. : 1. / norm; v[0] *= invnorm; v[1] *= invnorm; v[2] *= invnorm; }
      M->rot_omega = 2. * M_PI / atof(argv[10]);
      M->moving = 2;
    } else return fail(e, "supported: 'linear vx vy vz' | 'rotate origin x y z axis x y z period T'");
    M->next_reneighbor = -1; return 0; }
  return fail(e, "no such mesh");
}
int orc_add_wall_mesh(orc_engine *e, const char *id, int argc, const char *const *argv)
{ /* fix wall/gran model ... mesh n_meshes N meshes id... : fix_wall_gran.cpp:171-330 */
  if (e->nmwalls == MAXMESH) return fail(e, "too many mesh walls");
  meshwall_t *W = &e->mwalls[e->nmwalls]; memset(W, 0, sizeof *W); snprintf(W->id, sizeof W->id, "%s", id);
  if (parse_model(e, &argc, &argv, &W->m)) return -1;
  if (argc < 4 || strcmp(argv[0], "mesh") || strcmp(argv[1], "n_meshes")) return fail(e, "expected 'mesh n_meshes N meshes ...'");
  W->nmesh = atoi(argv[2]);
  if (W->nmesh < 1 || W->nmesh > MAXMESH || argc < 4 + W->nmesh || strcmp(argv[3], "meshes")) return fail(e, "bad mesh list");
  for (int k = 0; k < W->nmesh; k++) {
    int found = -1; for (int m = 0; m < e->nmeshes; m++) if (!strcmp(e->meshes[m].id, argv[4 + k])) found = m;
    if (found < 0) return fail(e, "unknown mesh id");
    W->mesh[k] = found; e->meshes[found].wall = e->nmwalls;
  }
  argv += 4 + W->nmesh; argc -= 4 + W->nmesh;
  if (parse_settings(e, argc, argv, &W->m)) return -1;
  e->nmwalls++; return 0;
}

/* TriMesh::resolveTriSphereNeighbuild tri_mesh_I.h:275-305 */
static int tri_sphere_neighbuild(const mesh_t *M, int t, double rSphere, const double *c, double treshold)
{
  const double maxDist = rSphere + treshold;
  double v[3]; v3sub(c, M->center[t], v);
  if (fabs(v3dot(M->surfNorm[t], v)) > maxDist) return 0;
  const double dParaMax = maxDist * maxDist;
  for (int i = 0; i < 3; i++) { v3sub(c, M->node[t][i], v); const double d = v3dot(M->edgeNorm[t][i], v); if (d > 0 && d * d > dParaMax) return 0; }
  return 1;
}
static double calc_dist(const double *cs, const double *cp, double *delta)
{ v3sub(cp, cs, delta); return sqrt((cs[0] - cp[0]) * (cs[0] - cp[0]) + (cs[1] - cp[1]) * (cs[1] - cp[1]) + (cs[2] - cp[2]) * (cs[2] - cp[2])); } /* tri_mesh_I.h:309-313 */
static double resolve_edge(const mesh_t *M, int t, int iEdge, const double *p, double *delta, double *bary)
{ /* tri_mesh_I.h:129-170, skip_inactive = true */
  const int ip = (iEdge + 1) % 3, ipp = (iEdge + 2) % 3; double nodeToP[3];
  v3sub(p, M->node[t][iEdge], nodeToP);
  const double distFromNode = v3dot(nodeToP, M->edgeVec[t][iEdge]);
  if (distFromNode < -SMALL_TRIMESH) {
    if (!M->cornerActive[t][iEdge]) return LARGE_TRIMESH;
    bary[iEdge] = 1.; bary[ip] = 0.; bary[ipp] = 0.;
    return calc_dist(p, M->node[t][iEdge], delta);
  } else if (distFromNode > M->edgeLen[t][iEdge] + SMALL_TRIMESH) {
    if (!M->cornerActive[t][ip]) return LARGE_TRIMESH;
    bary[iEdge] = 0.; bary[ip] = 1.; bary[ipp] = 0.;
    return calc_dist(p, M->node[t][ip], delta);
  }
  if (!M->edgeActive[t][iEdge]) return LARGE_TRIMESH;
  double cp[3]; for (int d = 0; d < 3; d++) cp[d] = M->node[t][iEdge][d] + distFromNode * M->edgeVec[t][iEdge][d];
  const double dd = calc_dist(p, cp, delta);
  bary[ipp] = 0.; bary[iEdge] = 1. - distFromNode / M->edgeLen[t][iEdge]; bary[ip] = 1. - bary[iEdge];
  return dd;
}
static double resolve_corner(const mesh_t *M, int t, int iNode, int obtuse, const double *p, double *delta, double *bary)
{ /* tri_mesh_I.h:174-255 */
  const int ip = (iNode + 1) % 3, ipp = (iNode + 2) % 3; const double *n = M->node[t][iNode];
  if (obtuse) {
    double nodeToP[3], cp[3]; v3sub(p, n, nodeToP);
    double distFromNode = v3dot(nodeToP, M->edgeVec[t][ipp]);
    if (distFromNode < SMALL_TRIMESH) {
      if (distFromNode > -M->edgeLen[t][ipp]) {
        if (!M->edgeActive[t][ipp]) return LARGE_TRIMESH;
        for (int d = 0; d < 3; d++) cp[d] = n[d] + distFromNode * M->edgeVec[t][ipp][d];
        bary[ip] = 0.; bary[iNode] = 1. + distFromNode / M->edgeLen[t][ipp]; bary[ipp] = 1. - bary[iNode];
        return calc_dist(p, cp, delta);
      } else {
        if (!M->cornerActive[t][ipp]) return LARGE_TRIMESH;
        bary[ipp] = 1.; bary[iNode] = bary[ip] = 0.;
        return calc_dist(p, M->node[t][ipp], delta);
      }
    }
    distFromNode = v3dot(nodeToP, M->edgeVec[t][iNode]);
    if (distFromNode > -SMALL_TRIMESH) {
      if (distFromNode < M->edgeLen[t][iNode]) {
        if (!M->edgeActive[t][iNode]) return LARGE_TRIMESH;
        for (int d = 0; d < 3; d++) cp[d] = n[d] + distFromNode * M->edgeVec[t][iNode][d];
        bary[ipp] = 0.; bary[iNode] = 1. - distFromNode / M->edgeLen[t][iNode]; bary[ip] = 1. - bary[iNode];
        return calc_dist(p, cp, delta);
      } else {
        if (!M->cornerActive[t][ip]) return LARGE_TRIMESH;
        bary[ip] = 1.; bary[iNode] = bary[ipp] = 0.;
        return calc_dist(p, M->node[t][ip], delta);
      }
    }
  }
  if (!M->cornerActive[t][iNode]) return LARGE_TRIMESH;
  bary[iNode] = 1.; bary[ip] = bary[ipp] = 0.;
  return calc_dist(p, n, delta);
}
/* TriMesh::resolveTriSphereContactBary tri_mesh_I.h:65-127 ; returns distance - radius */
static double tri_sphere_contact(const mesh_t *M, int t, double rSphere, const double *c, double *delta, double *bary, int *barySign)
{
  double n0c[3]; v3sub(c, M->node[t][0], n0c);
  bary[0] = bary[1] = bary[2] = 0.;
  { /* MathExtraLiggghts::calcBaryTriCoords math_extra_liggghts.h:583-594 */
    const double a = v3dot(n0c, M->edgeVec[t][0]), b = v3dot(n0c, M->edgeVec[t][2]), cc = v3dot(M->edgeVec[t][0], M->edgeVec[t][2]);
    const double oneMinCSqr = 1 - cc * cc;
    bary[1] = (a - b * cc) / (M->edgeLen[t][0] * oneMinCSqr);
    bary[2] = (a * cc - b) / (M->edgeLen[t][2] * oneMinCSqr);
    bary[0] = 1. - bary[1] - bary[2];
  }
  const double invlen = 1. / (2. * M->rbound[t]);
  const int bs = (bary[0] > -M->precision * invlen) + 2 * (bary[1] > -M->precision * invlen) + 4 * (bary[2] > -M->precision * invlen);
  *barySign = bs;
  const int ob = M->obtuse[t];
  double d = 1.;
  switch (bs) {
    case 1: d = resolve_corner(M, t, 0, ob == 0, c, delta, bary); break;
    case 2: d = resolve_corner(M, t, 1, ob == 1, c, delta, bary); break;
    case 3: d = resolve_edge(M, t, 0, c, delta, bary); break;
    case 4: d = resolve_corner(M, t, 2, ob == 2, c, delta, bary); break;
    case 5: d = resolve_edge(M, t, 2, c, delta, bary); break;
    case 6: d = resolve_edge(M, t, 1, c, delta, bary); break;
    case 7: { /* resolveFaceContactBary :259-271 */
      const double dNorm = v3dot(M->surfNorm[t], n0c); double cs[3];
      for (int k = 0; k < 3; k++) cs[k] = c[k] - M->surfNorm[t][k] * dNorm;
      d = calc_dist(c, cs, delta); break; }
    default: d = 1.; break;
  }
  return d - rSphere;
}

static double cutneighmax_of(const orc_engine *e)
{ double rmax = 0.0; for (long i = 0; i < e->n; i++) if (e->radius[i] > rmax) rmax = e->radius[i]; return 2.0 * rmax * e->cdf + e->skin; }

static void mesh_sort_contacts(orc_engine *e, mesh_t *M)
{ /* FixContactHistoryMesh::sort_contacts fix_contact_history_mesh.cpp:375-403 (pre_exchange) */
  const int dnum = e->mwalls[M->wall].m.dnum;
  if (!M->nneighs) return;
  for (long i = 0; i < e->n; i++) {
    const int nn = M->nneighs[i]; if (!nn) continue;
    int fe, lf;
    do {
      fe = lf = -1;
      for (int j = 0; j < nn; j++) { if (fe == -1 && M->partner[i][j] == -1) fe = j; if (M->partner[i][j] >= 0) lf = j; }
      if (fe > -1 && lf > -1 && fe < lf) { /* swap(i, fe, lf, true) */
        int t = M->partner[i][fe]; M->partner[i][fe] = M->partner[i][lf]; M->partner[i][lf] = t;
        for (int d = 0; d < dnum; d++) { double h = M->chist[i][fe * dnum + d]; M->chist[i][fe * dnum + d] = M->chist[i][lf * dnum + d]; M->chist[i][lf * dnum + d] = h; }
      }
    } while (fe > -1 && lf > -1 && fe < lf);
  }
}
static int contact_in_list(const mesh_t *M, int t, int i) { for (int k = 0; k < M->ncontacts[t]; k++) if (M->contacts[t][k] == i) return 1; return 0; }

static void mesh_build(orc_engine *e, mesh_t *M)
{ /* FixNeighlistMesh::pre_force fix_neighlist_mesh.cpp:230-307 + FixContactHistoryMesh::pre_force fix_contact_history_mesh.cpp:315-371 */
  const long n = e->n; const int dnum = e->mwalls[M->wall].m.dnum;
  const double skin = M->moving ? e->skin : 0.5 * e->skin;
  int *nn_new = (int *)calloc(n ? n : 1, sizeof(int));
  for (int t = 0; t < M->ntri; t++) {
    M->ncontacts[t] = 0;
    for (long i = 0; i < n; i++)
      if (tri_sphere_neighbuild(M, t, e->radius[i] * e->cdf, &e->x[3 * i], skin)) {
        if (M->ncontacts[t] == M->capcontacts[t]) { M->capcontacts[t] = M->capcontacts[t] ? 2 * M->capcontacts[t] : 8; M->contacts[t] = realloc(M->contacts[t], sizeof(int) * M->capcontacts[t]); }
        M->contacts[t][M->ncontacts[t]++] = (int)i; nn_new[i]++;
      }
  }
  if (!M->nneighs) { /* first build */
    M->nneighs = (int *)calloc(n ? n : 1, sizeof(int)); M->npartner = (int *)calloc(n ? n : 1, sizeof(int));
    M->partner = (int **)calloc(n ? n : 1, sizeof(int *)); M->chist = (double **)calloc(n ? n : 1, sizeof(double *)); M->keep = (unsigned char **)calloc(n ? n : 1, sizeof(unsigned char *));
  }
  for (long i = 0; i < n; i++) {
    /* cleanUpContactJumps :467-504 */
    int ip = 0;
    while (ip < M->npartner[i]) {
      if (!contact_in_list(M, M->partner[i][ip], (int)i)) {
        const int last = M->npartner[i] - 1;
        M->partner[i][ip] = -1; for (int d = 0; d < dnum; d++) M->chist[i][ip * dnum + d] = 0.0;
        int t = M->partner[i][ip]; M->partner[i][ip] = M->partner[i][last]; M->partner[i][last] = t;
        for (int d = 0; d < dnum; d++) { double h = M->chist[i][ip * dnum + d]; M->chist[i][ip * dnum + d] = M->chist[i][last * dnum + d]; M->chist[i][last * dnum + d] = h; }
        M->npartner[i]--;
      } else ip++;
    }
    const int nn = nn_new[i];
    int *pn = (int *)malloc(sizeof(int) * (nn ? nn : 1)); double *hn = (double *)calloc((size_t)(nn ? nn : 1) * (dnum ? dnum : 1), sizeof(double));
    for (int k = 0; k < nn; k++) pn[k] = -1;
    for (int k = 0; k < M->npartner[i] && k < nn; k++) { pn[k] = M->partner[i][k]; for (int d = 0; d < dnum; d++) hn[k * dnum + d] = M->chist[i][k * dnum + d]; }
    free(M->partner[i]); free(M->chist[i]); free(M->keep[i]);
    M->partner[i] = pn; M->chist[i] = hn; M->keep[i] = (unsigned char *)calloc(nn ? nn : 1, 1); M->nneighs[i] = nn;
  }
  free(nn_new);
  memcpy(M->nodesLastRe, M->node, sizeof(double) * 9 * M->ntri); /* storeNodePosRebuild */
}

static void mesh_wall_compute(orc_engine *e, meshwall_t *W, int shearupdate)
{ /* FixWallGran::post_force_mesh fix_wall_gran.cpp:803-982 + fix_contact_history_mesh_I.h:51-215 */
  const int dnum = W->m.dnum; const double cdmul = e->cdf - 1.0; const double cutneighmax = cutneighmax_of(e);
  for (int im = 0; im < W->nmesh; im++) {
    mesh_t *M = &e->meshes[W->mesh[im]];
    for (int d = 0; d < 3; d++) M->f_total[d] = M->torque_total[d] = 0.; /* MeshModuleStress::pre_force :286-297 */
    for (long i = 0; i < e->n; i++) for (int k = 0; k < M->nneighs[i]; k++) M->keep[i][k] = 0; /* markAllContacts */
    for (int t = 0; t < M->ntri; t++) for (int c = 0; c < M->ncontacts[t]; c++) {
      const int ip = M->contacts[t][c];
      double delta[3], bary[3]; int barysign = -1;
      const double radi = e->radius[ip];
      const double deltan = tri_sphere_contact(M, t, radi, &e->x[3 * ip], delta, bary, &barysign);
      if (deltan > cutneighmax) continue;
      const int intersect = (deltan <= 0);
      if (!(deltan <= 0 || deltan < cdmul * radi)) continue;
      /* handleContact */
      const int nn = M->nneighs[ip]; int *tri = M->partner[ip]; double *hist = NULL; int have = 0;
      for (int k = 0; k < nn; k++) if (tri[k] == t) { hist = &M->chist[ip][k * dnum]; M->keep[ip][k] = 1; have = 1; break; }
      if (!have) {
        const int faceflag = (7 == barysign);
        if (faceflag) { /* coplanarContactAlready */
          int already = 0;
          for (int k = 0; k < nn; k++) { const int q = tri[k]; if (q >= 0 && q != t && coplanar_node_neighs(M, q, t) && M->keep[ip][k]) { already = 1; break; } }
          if (already) continue;
        }
        int ic = -1; for (int k = 0; k < nn; k++) if (tri[k] == -1) { ic = k; break; } /* addNewTriContactToExistingParticle */
        if (ic < 0) { snprintf(e->err, sizeof e->err, "mesh contact rows full"); continue; }
        tri[ic] = t; M->keep[ip][ic] = 1; hist = &M->chist[ip][ic * dnum];
        for (int d = 0; d < dnum; d++) hist[d] = 0.0;
        M->npartner[ip]++;
        if (faceflag) for (int k = 0; k < nn; k++) if (tri[k] >= 0 && tri[k] != t && coplanar_node_neighs(M, tri[k], t)) for (int d = 0; d < dnum; d++) hist[d] = M->chist[ip][k * dnum + d]; /* checkCoplanarContactHistory */
      }
      double v_wall[3] = {0., 0., 0.};
      if (M->moving) for (int d = 0; d < 3; d++) v_wall[d] = (bary[0] * M->vnode[t][0][d] + bary[1] * M->vnode[t][1][d] + bary[2] * M->vnode[t][2][d]);
      if (intersect) { /* Walls::Granular::compute_force fix_wall_gran_base.h:159-367 */
        sid_t s_; sid_t *sd = &s_; memset(sd, 0, sizeof *sd);
        sd->is_wall = 1; sd->radi = radi; sd->deltan = -deltan;
        sd->delta[0] = -delta[0]; sd->delta[1] = -delta[1]; sd->delta[2] = -delta[2];
        sd->vi = &e->v[3 * ip]; sd->vj = v_wall; sd->wi = &e->omega[3 * ip]; sd->wj = NULL;
        sd->r = sd->radi - sd->deltan; sd->rinv = 1.0 / sd->r;
        sd->itype = e->type[ip]; sd->jtype = M->atom_type; sd->meff = e->rmass[ip]; sd->mi = e->rmass[ip];
        sd->shearupdate = shearupdate; sd->radsum = sd->radi;
        for (int d = 0; d < 3; d++) sd->en[d] = sd->delta[d] * sd->rinv;
        sd->hist = hist; sd->flag = NULL;
        chain_intersect(e, &W->m, sd);
        double force_old[3] = {e->f[3 * ip], e->f[3 * ip + 1], e->f[3 * ip + 2]};
        for (int d = 0; d < 3; d++) { e->f[3 * ip + d] += sd->Fi[d]; e->torque[3 * ip + d] += sd->Ti[d]; }
        if (M->stress) { /* fix_wall_gran_base.h:350-362 (f_pw = f - force_old) + MeshModuleStress::add_particle_contribution :316-345 */
          double frc[3], cp[3], tmp[3];
          for (int d = 0; d < 3; d++) { frc[d] = -(e->f[3 * ip + d] - force_old[d]); cp[d] = e->x[3 * ip + d] + delta[d]; }
          for (int d = 0; d < 3; d++) { M->f_total[d] = M->f_total[d] + frc[d]; tmp[d] = cp[d] - M->p_ref[d]; }
          M->torque_total[0] = M->torque_total[0] + (tmp[1] * frc[2] - tmp[2] * frc[1]);
          M->torque_total[1] = M->torque_total[1] + (tmp[2] * frc[0] - tmp[0] * frc[2]);
          M->torque_total[2] = M->torque_total[2] + (tmp[0] * frc[1] - tmp[1] * frc[0]);
        }
      } else chain_close(&W->m, hist, NULL);
    }
    /* cleanUpContacts :437-463 */
    for (long i = 0; i < e->n; i++) for (int k = 0; k < M->nneighs[i]; k++) if (!M->keep[i][k]) {
      if (M->partner[i][k] > -1) M->npartner[i]--;
      M->partner[i][k] = -1; for (int d = 0; d < dnum; d++) M->chist[i][k * dnum + d] = 0.0;
    }
  }
}

/* MathExtra::quatquat math_extra.h:596-602 ; MathExtraLiggghts::vec_quat_rotate math_extra_liggghts.h:435-474 */
static void quatquat(const double *a, const double *b, double *c)
{
  c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  c[1] = a[0] * b[1] + b[0] * a[1] + a[2] * b[3] - a[3] * b[2];
  c[2] = a[0] * b[2] + b[0] * a[2] + a[3] * b[1] - a[1] * b[3];
  c[3] = a[0] * b[3] + b[0] * a[3] + a[1] * b[2] - a[2] * b[1];
}
static void vec_quat_rotate(double *vec, const double *quat)
{
  double vecQ[4] = {0., vec[0], vec[1], vec[2]}, quatC[4] = {quat[0], -quat[1], -quat[2], -quat[3]}, temp[4], resultQ[4];
  quatquat(quat, vecQ, temp); quatquat(temp, quatC, resultQ);
  vec[0] = resultQ[1]; vec[1] = resultQ[2]; vec[2] = resultQ[3];
}
static void mesh_rotate_step(orc_engine *e, mesh_t *M)
{ /* MeshMoverRotate::initial_integrate mesh_mover_rotation.cpp:98-125 -> FixMesh::rotate fix_mesh.cpp:747-759 ->
   * MultiNodeMesh::rotate(dAngle,axis,p) multi_node_mesh_I.h:620-672 + TrackingMesh::rotate tracking_mesh_I.h:400-411
   * (edgeVec, edgeNorm, surfaceNorm are the rotating element properties, surface_mesh_I.h:78-80; GeneralContainer::rotate
   * general_container_I.h:642-651); then v_node = 0 + omegaVec x (node - reference point) */
  const double dphi = M->rot_omega * e->dt;
  double axisNorm[3] = {M->rot_axis[0], M->rot_axis[1], M->rot_axis[2]};
  { const double sinv = 1. / sqrt(axisNorm[0] * axisNorm[0] + axisNorm[1] * axisNorm[1] + axisNorm[2] * axisNorm[2]);
    axisNorm[0] = sinv * axisNorm[0]; axisNorm[1] = sinv * axisNorm[1]; axisNorm[2] = sinv * axisNorm[2]; }
  double dQ[4]; dQ[0] = cos(dphi * 0.5); for (int i = 0; i < 3; i++) dQ[i + 1] = axisNorm[i] * sin(dphi * 0.5);
  const double *origin = M->rot_origin;
  const int trans = (origin[0] * origin[0] + origin[1] * origin[1] + origin[2] * origin[2]) > 0.;
  double omegaVec[3]; for (int d = 0; d < 3; d++) omegaVec[d] = M->rot_axis[d] * M->rot_omega;
  for (int t = 0; t < M->ntri; t++) {
    double *c = M->center[t]; c[0] = c[1] = c[2] = 0.;
    for (int j = 0; j < 3; j++) { double *nd = M->node[t][j];
      if (trans) for (int d = 0; d < 3; d++) nd[d] = nd[d] - origin[d];
      vec_quat_rotate(nd, dQ);
      if (trans) for (int d = 0; d < 3; d++) nd[d] = nd[d] + origin[d];
      for (int d = 0; d < 3; d++) c[d] = nd[d] + c[d]; }
    { const double sinv = 1. / 3.; c[0] = sinv * c[0]; c[1] = sinv * c[1]; c[2] = sinv * c[2]; }
    for (int j = 0; j < 3; j++) { vec_quat_rotate(M->edgeVec[t][j], dQ); vec_quat_rotate(M->edgeNorm[t][j], dQ); }
    vec_quat_rotate(M->surfNorm[t], dQ);
    for (int j = 0; j < 3; j++) { double rPA[3], vRot[3]; v3sub(M->node[t][j], origin, rPA);
      vRot[0] = omegaVec[1] * rPA[2] - omegaVec[2] * rPA[1]; vRot[1] = omegaVec[2] * rPA[0] - omegaVec[0] * rPA[2]; vRot[2] = omegaVec[0] * rPA[1] - omegaVec[1] * rPA[0];
      for (int d = 0; d < 3; d++) M->vnode[t][j][d] = 0. + vRot[d]; }
  }
  { double *q = M->p_ref; /* TrackingMesh::rotate: customValues_.move(-origin), rotate(dQ), move(origin) on the global properties too */
    if (trans) for (int d = 0; d < 3; d++) q[d] = q[d] + (-origin[d]);
    vec_quat_rotate(q, dQ);
    if (trans) for (int d = 0; d < 3; d++) q[d] = q[d] + origin[d]; }
}
static void mesh_move_step(orc_engine *e)
{ /* FixMoveMesh::initial_integrate fix_move_mesh.cpp:221-238 + MeshMoverLinear::initial_integrate mesh_mover_linear.cpp:94-112
   * + MultiNodeMesh::move(vecIncremental) multi_node_mesh_I.h:502-526 */
  for (int m = 0; m < e->nmeshes; m++) { mesh_t *M = &e->meshes[m]; if (!M->moving) continue;
    if (M->moving == 2) { mesh_rotate_step(e, M); continue; }
    double dx[3]; for (int d = 0; d < 3; d++) dx[d] = M->vel[d] * e->dt;
    for (int t = 0; t < M->ntri; t++) {
      for (int j = 0; j < 3; j++) for (int d = 0; d < 3; d++) { M->node[t][j][d] = M->node[t][j][d] + dx[d]; M->vnode[t][j][d] = 0. + M->vel[d]; }
      for (int d = 0; d < 3; d++) M->center[t][d] = M->center[t][d] + dx[d];
    }
    for (int d = 0; d < 3; d++) M->p_ref[d] = M->p_ref[d] + dx[d]; /* p_ref is a frame_general mesh property: it travels with the mesh (mesh_module_stress.cpp:75-80) */
  }
}
static int mesh_decide_rebuild(orc_engine *e)
{ /* FixMesh::pre_force fix_mesh.cpp:553-576 + MultiNodeMesh::decideRebuild multi_node_mesh_I.h:792-826: a node moved more than
   * skin/2 since the last build -> next_reneighbor = ntimestep + 1 */
  int any = 0;
  for (int m = 0; m < e->nmeshes; m++) { mesh_t *M = &e->meshes[m]; if (!M->moving) continue;
    const double trig = 0.25 * e->skin * e->skin; int flag = 0;
    for (int t = 0; t < M->ntri && !flag; t++) for (int j = 0; j < 3; j++) { double dd[3]; v3sub(M->node[t][j], M->nodesLastRe[t][j], dd); if (dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2] > trig) flag = 1; }
    if (flag) { M->next_reneighbor = (int)e->ntimestep + 1; any = 1; } }
  return any;
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
  for (int m = 0; m < e->nmeshes; m++) if (e->meshes[m].wall >= 0) mesh_sort_contacts(e, &e->meshes[m]); /* pre_exchange */
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
  for (int m = 0; m < e->nmeshes; m++) if (e->meshes[m].moving) for (int t = 0; t < e->meshes[m].ntri; t++) tri_surf_properties(&e->meshes[m], t); /* refreshOwned */
  for (int m = 0; m < e->nmeshes; m++) if (e->meshes[m].wall >= 0) mesh_build(e, &e->meshes[m]);
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
        sd->rsq = rsq; sd->ntimestep = e->ntimestep; sd->xi = &e->x[3 * i];
        chain_intersect(e, &e->pm, sd);
        for (int d = 0; d < 3; d++) { e->f[3 * i + d] += sd->Fi[d]; e->torque[3 * i + d] += sd->Ti[d]; e->f[3 * j + d] += sd->Fj[d]; e->torque[3 * j + d] += sd->Tj[d]; }
      } else if (rsq < cdm * radsum * radsum) { /* :420 ; unreachable when cdf == 1 */
        if (e->pm.cohesion) { /* ContactModel::surfacesClose contact_models.h:246-253: the bond may act across the gap */
          sid_t s_; sid_t *sd = &s_; memset(sd, 0, sizeof *sd);
          sd->itype = e->type[i]; sd->jtype = e->type[j]; sd->shearupdate = shearupdate; sd->radi = radi; sd->radj = radj; sd->radsum = radsum;
          sd->delta[0] = delx; sd->delta[1] = dely; sd->delta[2] = delz; sd->rsq = rsq; sd->ntimestep = e->ntimestep; sd->xi = &e->x[3 * i];
          sd->vi = &e->v[3 * i]; sd->vj = &e->v[3 * j]; sd->wi = &e->omega[3 * i]; sd->wj = &e->omega[3 * j];
          sd->hist = &e->hist[m * dnum]; sd->flag = &e->flag[m];
          cohesion_bond(e, &e->pm, sd);
          if (sd->has_force_update) for (int d = 0; d < 3; d++) { e->f[3 * i + d] += sd->Fi[d]; e->torque[3 * i + d] += sd->Ti[d]; e->f[3 * j + d] += sd->Fj[d]; e->torque[3 * j + d] += sd->Tj[d]; }
        }
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
  for (int w = 0; w < e->nmwalls; w++) mesh_wall_compute(e, &e->mwalls[w], shearupdate);
  for (int q = 0; q < e->nxf; q++) for (long i = 0; i < e->n; i++) if (e->mask[i] & e->xf[q].bit) {
    if (e->xf[q].kind == 0) for (int d = 0; d < 3; d++) e->f[3 * i + d] += e->xf[q].v[d]; /* fix_addforce.cpp:234-260 (constant components) */
    else { const double drag = e->xf[q].v[0]; for (int d = 0; d < 3; d++) e->f[3 * i + d] -= drag * e->v[3 * i + d]; } /* fix_viscous.cpp:100-125 */
  }
  if (e->freezebit) for (long i = 0; i < e->n; i++) if (e->mask[i] & e->freezebit) for (int d = 0; d < 3; d++) { e->f[3 * i + d] = 0.0; e->torque[3 * i + d] = 0.0; } /* fix_freeze.cpp:132-144 */
}

static int setup_prepare(orc_engine *e, int first)
{
  if (!e->n && !e->tag) return fail(e, "no particles uploaded");
  derive_tables(e);
  e->cdf = e->cdf_user > 1.0 ? e->cdf_user : 1.0;
  if (e->have_pair && e->pm.cohesion) { /* cohesion_model_bond.h:391-475 / cohesion_model_bond_nonlinear.h:336-382: neighbor->register_contact_dist_factor */
    double minrad = 1e99; for (long i = 0; i < e->n; i++) if (e->radius[i] < minrad) minrad = e->radius[i];
    double cdf_all = e->cdf; /* register_contact_dist_factor: max(existing, requested), neighbor.h:142 */
    for (int i = 1; i <= e->ntypes; i++) for (int j = 1; j <= e->ntypes; j++) {
      double one;
      if (!e->pm.stressBreak) one = 1.1 * 0.5 * e->bp[BP_MAXDIST][i][j] / minrad;
      else {
        double stress = e->bp[BP_MAXSIGMA][i][j];
        if (e->pm.ratioTC) stress = fmax(stress, stress * e->bp[BP_RATIOTC][i][j]);
        if (e->pm.cohesion == C_BOND) one = 0.5 * (1.1 * e->bp[BP_CREATEDIST][i][j] / minrad + 1.1 * stress / (e->bp[BP_KN][i][j] * minrad));
        else one = 0.5 * 1.1 * e->bp[BP_CREATEDIST][i][j] / minrad + 0.5 * 1.1 * stress / (e->bp[BP_K_FN2][i][j] * minrad);
      }
      cdf_all = one > cdf_all ? one : cdf_all;
    }
    if (cdf_all > 10.) return fail(e, "Maximum bond distance exceeding 10 x particle diameter");
    e->cdf = cdf_all;
  }
  for (int w = 0; w < e->nwalls; w++) if (!e->walls[w].hist) { e->walls[w].hist = (double *)calloc((size_t)(e->n ? e->n : 1) * (e->walls[w].m.dnum ? e->walls[w].m.dnum : 1), sizeof(double)); }
  if (first) for (int m = 0; m < e->nmeshes; m++) if (e->meshes[m].moving) memset(e->meshes[m].vnode, 0, sizeof(double) * 9 * e->meshes[m].ntri); /* FixMoveMesh::setup fix_move_mesh.cpp:194-217: v = 0 */
  return 0;
}

int orc_setup(orc_engine *e)
{ /* Verlet::setup verlet.cpp:134-199 */
  if (e->ins_open) return fail(e, "setup inside an insertion step");
  const int rc = setup_prepare(e, 1); if (rc) return rc;
  build(e);
  e->nbuilds = 0; /* neighbor->ncalls counts builds of the current run only (neighbor.cpp init: ncalls = 0) */
  compute_forces(e, 0);
  e->setup_done = 1; return 0;
}

static void first_half_step(orc_engine *e)
{
  const double dtv = e->dt, dtf = 0.5 * e->dt * e->ftm2v, dtfrotate = dtf / 0.4; /* fix_nve.cpp:86, fix_nve_sphere.cpp:69,150 */
  for (long i = 0; i < e->n; i++) if (e->mask[i] & e->integbit) { /* fix_nve_sphere.cpp:134-183 */
    const double dtfm = dtf / (e->rmass[i] * (1. + 0.0 / e->density[i]));
    for (int d = 0; d < 3; d++) { e->v[3 * i + d] += dtfm * e->f[3 * i + d]; e->x[3 * i + d] += dtv * e->v[3 * i + d]; }
    const double dtirotate = dtfrotate / (e->radius[i] * e->radius[i] * e->rmass[i]);
    for (int d = 0; d < 3; d++) e->omega[3 * i + d] += dtirotate * e->torque[3 * i + d];
  }
}
static void second_half_step(orc_engine *e)
{
  const double dtf = 0.5 * e->dt * e->ftm2v, dtfrotate = dtf / 0.4;
  for (long i = 0; i < e->n; i++) if (e->mask[i] & e->integbit) { /* fix_nve_sphere.cpp:205-244 */
    const double dtfm = dtf / (e->rmass[i] * (1. + 0.0 / e->density[i]));
    for (int d = 0; d < 3; d++) e->v[3 * i + d] += dtfm * e->f[3 * i + d];
    const double dtirotate = dtfrotate / (e->radius[i] * e->radius[i] * e->rmass[i]);
    for (int d = 0; d < 3; d++) e->omega[3 * i + d] += dtirotate * e->torque[3 * i + d];
  }
}

/* The timestep in which fix insert/* creates particles (FixInsert::pre_exchange fix_insert.cpp:672-905 sits between
 * initial_integrate and the rebuild it forces, verlet.cpp:277-310), in two halves so that the caller can look at the positions
 * the overlap check of the reference sees. */
int orc_insert_step_begin(orc_engine *e)
{
  if (e->ins_open) return fail(e, "insert_step_begin called twice");
  if (e->tag) {
    if (!e->setup_done) return fail(e, "insert_step_begin before setup");
    e->ntimestep++;
    first_half_step(e);
    mesh_move_step(e);
  } else e->ntimestep++;
  e->ins_open = 1; return 0;
}
int orc_insert_step_end(orc_engine *e, long n, const int *tag, const int *type, const int *mask,
                        const double *x, const double *v, const double *omega, const double *radius, const double *density)
{
  if (!e->ins_open) return fail(e, "insert_step_end without insert_step_begin");
  if (!e->tag && n <= 0) { e->ins_open = 0; return 0; } /* a step of an empty box */
  const int first = !e->setup_done;
  e->ins_mass = 1;
  const int rc = n > 0 ? orc_insert_particles(e, n, tag, type, mask, x, v, omega, radius, density) : 0;
  e->ins_mass = 0;
  if (rc) return rc;
  const int rc2 = setup_prepare(e, first); if (rc2) return rc2;
  if (first) e->nbuilds = 0; /* (the run of an empty box had no setup that would have reset neighbor->ncalls) */
  build(e); /* fix->next_reneighbor forces the rebuild, neighbor.cpp:1364-1369 */
  compute_forces(e, 1);
  second_half_step(e);
  e->ins_open = 0; e->setup_done = 1; return 0;
}

int orc_run(orc_engine *e, long nsteps)
{ /* Verlet::run verlet.cpp:264-391 */
  if (!e->setup_done) return fail(e, "run before setup");
  if (e->ins_open) return fail(e, "run inside an insertion step");
  for (long s = 0; s < nsteps; s++) {
    e->ntimestep++;
    first_half_step(e);
    mesh_move_step(e);
    /* Neighbor::decide neighbor.cpp:1362-1376 + check_distance :1425-1466 */
    int nflag = 0, forced = 0;
    for (int m = 0; m < e->nmeshes; m++) if (e->meshes[m].moving && e->meshes[m].next_reneighbor == (int)e->ntimestep) forced = 1; /* fix->next_reneighbor, neighbor.cpp:1364-1369 */
    if (forced) nflag = 1; else e->ago++;
    if (!forced && e->ago >= e->delay && e->ago % e->every == 0) {
      if (!e->check) nflag = 1;
      else { const double deltasq = 0.25 * e->skin * e->skin;
        for (long i = 0; i < e->n; i++) { const double dx = e->x[3 * i] - e->xhold[3 * i], dy = e->x[3 * i + 1] - e->xhold[3 * i + 1], dz = e->x[3 * i + 2] - e->xhold[3 * i + 2];
          if (dx * dx + dy * dy + dz * dz > deltasq) nflag = 1; } }
    }
    if (nflag) build(e); else mesh_decide_rebuild(e);
    compute_forces(e, 1);
    second_half_step(e);
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
    if (hist) for (int d = 0; d < dnum; d++) hist[r * dnum + d] = (rows[r].swap && e->pm.nflag[d]) ? -e->hist[rows[r].m * dnum + d] : e->hist[rows[r].m * dnum + d]; }
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
/* mesh read-back: topology (for pinning) and per-particle contact rows sorted by (tag, triangle) */
int orc_download_mesh(orc_engine *e, const char *mesh_id, const char *field, void *out, long count)
{
  for (int m = 0; m < e->nmeshes; m++) if (!strcmp(e->meshes[m].id, mesh_id)) { mesh_t *M = &e->meshes[m]; const int T = M->ntri;
    if (!strcmp(field, "nodes") && count == 9L * T) { memcpy(out, M->node, sizeof(double) * 9 * T); return 0; }
    if (!strcmp(field, "edge_vec") && count == 9L * T) { memcpy(out, M->edgeVec, sizeof(double) * 9 * T); return 0; }
    if (!strcmp(field, "edge_norm") && count == 9L * T) { memcpy(out, M->edgeNorm, sizeof(double) * 9 * T); return 0; }
    if (!strcmp(field, "surf_norm") && count == 3L * T) { memcpy(out, M->surfNorm, sizeof(double) * 3 * T); return 0; }
    if (!strcmp(field, "center") && count == 3L * T) { memcpy(out, M->center, sizeof(double) * 3 * T); return 0; }
    if (!strcmp(field, "v_node") && count == 9L * T) { memcpy(out, M->vnode, sizeof(double) * 9 * T); return 0; }
    if (!strcmp(field, "edge_active") && count == 3L * T) { for (int k = 0; k < 3 * T; k++) ((int *)out)[k] = M->edgeActive[k / 3][k % 3]; return 0; }
    if (!strcmp(field, "corner_active") && count == 3L * T) { for (int k = 0; k < 3 * T; k++) ((int *)out)[k] = M->cornerActive[k / 3][k % 3]; return 0; }
    if (!strcmp(field, "obtuse") && count == T) { for (int k = 0; k < T; k++) ((int *)out)[k] = M->obtuse[k]; return 0; }
    if (!strcmp(field, "nneighs") && count == T) { for (int k = 0; k < T; k++) ((int *)out)[k] = M->nNeighs[k]; return 0; }
    return fail(e, "unknown mesh field or wrong count"); }
  return fail(e, "no such mesh");
}
int orc_mesh_force(orc_engine *e, const char *mesh_id, double *out9)
{ /* FixMeshSurface::compute_vector(0..8) of the stress module: total force, total torque about p_ref, p_ref */
  for (int m = 0; m < e->nmeshes; m++) if (!strcmp(e->meshes[m].id, mesh_id)) { mesh_t *M = &e->meshes[m];
    if (!M->stress) return fail(e, "mesh does not track stress (fix mesh/surface/stress)");
    for (int d = 0; d < 3; d++) { out9[d] = M->f_total[d]; out9[3 + d] = M->torque_total[d]; out9[6 + d] = M->p_ref[d]; }
    return 0; }
  return fail(e, "no such mesh");
}
int orc_mesh_contact_count(orc_engine *e, const char *mesh_id, long *n, int *dnum)
{
  for (int m = 0; m < e->nmeshes; m++) if (!strcmp(e->meshes[m].id, mesh_id)) { mesh_t *M = &e->meshes[m]; long c = 0;
    if (M->nneighs) for (long i = 0; i < e->n; i++) for (int k = 0; k < M->nneighs[i]; k++) c += M->partner[i][k] >= 0;
    *n = c; *dnum = M->wall >= 0 ? e->mwalls[M->wall].m.dnum : 0; return 0; }
  return fail(e, "no such mesh");
}
typedef struct { int tag, tri; const double *h; } mrow_t;
static int cmp_mrow(const void *a, const void *b) { const mrow_t *x = a, *y = b; if (x->tag != y->tag) return (x->tag > y->tag) - (x->tag < y->tag); return (x->tri > y->tri) - (x->tri < y->tri); }
int orc_download_mesh_contacts(orc_engine *e, const char *mesh_id, int *tag, int *tri, double *hist)
{
  for (int m = 0; m < e->nmeshes; m++) if (!strcmp(e->meshes[m].id, mesh_id)) { mesh_t *M = &e->meshes[m]; long c = 0; int dn = 0;
    orc_mesh_contact_count(e, mesh_id, &c, &dn);
    mrow_t *rows = (mrow_t *)malloc(sizeof(mrow_t) * (c ? c : 1)); long k2 = 0;
    if (M->nneighs) for (long i = 0; i < e->n; i++) for (int k = 0; k < M->nneighs[i]; k++) if (M->partner[i][k] >= 0) { rows[k2].tag = e->tag[i]; rows[k2].tri = M->partner[i][k]; rows[k2].h = &M->chist[i][k * dn]; k2++; }
    qsort(rows, k2, sizeof(mrow_t), cmp_mrow);
    for (long r = 0; r < k2; r++) { tag[r] = rows[r].tag; tri[r] = rows[r].tri; if (hist) for (int d = 0; d < dn; d++) hist[r * dn + d] = rows[r].h[d]; }
    free(rows); return 0; }
  return fail(e, "no such mesh");
}
typedef struct { long ntimestep, nbuilds, nlocal, nghost, npairs_full, ncontacts_full, kernel_launches; int maxneigh, dnum; double step_kernel_ms; long step_kernel_calls; } orc_stats;
/* compute bond/counter (compute_bond_counter.cpp:101-138) as lammps_extract_compute returns it between two runs: bonds created /
 * broken since the last call, and "total" = counted + created - broken in unsigned 32-bit arithmetic, where `counted` only
 * moves on the step right after an invocation DURING a run (bond_count, :158-168; Modify::init resets invoked_vector to -1 at
 * every run start) -- with invocations between runs it is 0.  The three wall entries stay 0 (no bond models on walls). */
int orc_bond_counter(orc_engine *e, double *out6)
{
  const unsigned int total = (unsigned int)e->bond_created - (unsigned int)e->bond_broken;
  out6[0] = (double)e->bond_created; out6[1] = (double)e->bond_broken; out6[2] = e->pm.cohesion == C_BOND ? (double)total : 0.0; out6[3] = out6[4] = out6[5] = 0.0;
  if (e->pm.cohesion != C_BOND) out6[0] = out6[1] = 0.0;  /* bond/nonlinear looks for a compute style that does not exist: never fed */
  e->bond_created = e->bond_broken = 0;
  return 0;
}
int orc_get_stats(orc_engine *e, orc_stats *s)
{ memset(s, 0, sizeof *s); s->ntimestep = e->ntimestep; s->nbuilds = e->nbuilds; s->nlocal = e->n; s->npairs_full = 2 * e->npairs; s->dnum = e->pm.dnum;
  long c = 0; for (long m = 0; m < e->npairs; m++) c += e->flag[m] != 0; s->ncontacts_full = 2 * c; return 0; }
