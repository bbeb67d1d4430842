// dem_engine.cu -- host orchestration + C ABI (include/dem_b200.h) of the B200 DEM engine.
// One dem_engine == one GPU.  The step loop mirrors Verlet::run (verlet.cpp:264-391) but the
// per-step work is ONE fused kernel (dem_kernels.cuh) plus a small ghost refresh; rebuilds
// run the sort / border / list kernels.  No CPU fallback exists anywhere in this file.
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dem_b200.h"
#include "dem_kernels.cuh"

using namespace dem;

#define MAXT 8

struct DemFail { int code; };
static void dem_fail(dem_engine *e, int code, const char *fmt, ...);

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess) dem_fail(E, DEM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  // grow to at least m elements; keep = number of leading elements to preserve
  void ensure(dem_engine *E, size_t m, size_t keep = 0, cudaStream_t st = 0);
};

struct ListSet {  // one ELLPACK neighbour list + history (two sets ping-pong across rebuilds)
  DevBuf<unsigned> nbr;
  DevBuf<int> ptag, numneigh;
  DevBuf<double4> hist;
  int cap = 0, maxk = 0, dnum = 0, hslots = 0, valid = 0;  // dnum = 32-byte history records per contact
};

struct WallHost { std::string id; WallP p; };

struct dem_engine {
  std::string err;
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = 0;
  // deck settings
  double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1}, prd[3] = {1, 1, 1};
  int periodic[3] = {0, 0, 0};
  int ntypes = 1;
  double skin = 0.0, dt = 0.0, nktv2p = 1.0, ftm2v = 1.0, cdf = 1.0;
  int every = 1, delay = 0, check = 1;
  double Y[MAXT + 1] = {0}, nu[MAXT + 1] = {0}, cor[MAXT + 1][MAXT + 1] = {{0}}, mu[MAXT + 1][MAXT + 1] = {{0}},
         rmu[MAXT + 1][MAXT + 1] = {{0}}, rvisc[MAXT + 1][MAXT + 1] = {{0}}, charVel = 0.0;
  std::map<std::string, int> have_prop;
  ModelP pm = {};
  int have_pair = 0;
  std::vector<WallHost> walls;
  int nwrows = 0;
  double g[3] = {0, 0, 0};
  int have_g = 0, freezebit = 0, integbit = 1;
  std::map<std::string, double> opt;
  // particles
  long nlocal = 0, nghost = 0;
  int cap = 0;
  double rmax = 0.0, cutneighmax = 0.0;
  DevBuf<double4> xr[2], vm[2], wt[2], xh;
  int cur = 0;
  DevBuf<int> tag, tag_tmp;
  DevBuf<double> density, density_tmp, f, tq, whist, whist_tmp, tab;
  double t1[T_COUNT] = {0};
  DevBuf<WallP> dwalls;
  DevBuf<unsigned> valid_tmp;
  DevBuf<int> wlist; DevBuf<double> fw; int nwc = 0, nwcap = 0;
  // ghosts
  DevBuf<int> gsrc, gshift, gsrc2, gshift2, flo, fhi, slo, shi;
  // cells / sort
  GridP grid = {};
  long ncells = 0;
  DevBuf<int> ocs, oce, gcs, gce, perm, vals;
  DevBuf<unsigned> keys, keys2;
  DevBuf<char> cubtmp;
  ListSet ls[2];
  int lcur = 0;
  DevBuf<int> overflow;
  DevBuf<unsigned long long> counters;
  int *hflag = nullptr;  // mapped pinned rebuild flag
  // state
  int uploaded = 0, setup_done = 0, forces_valid = 0;
  long ntimestep = 0, nbuilds = 0, launches = 0;
  int ago = 0;
  // timing of the step kernel
  std::vector<cudaEvent_t> ev;
  long ev_used = 0;
  double step_ms = 0.0;
  long step_calls = 0;
};

static void dem_fail(dem_engine *e, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (e) e->err = buf;
  throw DemFail{code};
}

template <typename T>
void DevBuf<T>::ensure(dem_engine *E, size_t m, size_t keep, cudaStream_t st)
{
  if (m <= n) return;
  T *q = nullptr;
  CK(cudaMalloc(&q, m * sizeof(T)));
  if (keep && p) CK(cudaMemcpyAsync(q, p, std::min(keep, n) * sizeof(T), cudaMemcpyDeviceToDevice, st));
  if (p) { CK(cudaStreamSynchronize(st)); cudaFree(p); }
  p = q; n = m;
}

#define GRID(n, b) (unsigned)(((n) + (b)-1) / (b))
#define API_BEGIN  if (!e) return DEM_ERR_ARG; dem_engine *E = e; (void)E; try {
#define API_END    } catch (const DemFail &f) { return f.code; } catch (const std::exception &x) { e->err = x.what(); return DEM_ERR_CUDA; } return DEM_OK;

// ------------------------------------------------------------------------------------------------
extern "C" const char *dem_version(void) { return "dem_b200 0.1 (sm_100a)"; }
extern "C" const char *dem_last_error(const dem_engine *e) { return e ? e->err.c_str() : "null engine"; }

extern "C" int dem_create(dem_engine **out, int device, int rank, int nranks, const void *nccl_id, void *stream)
{
  if (!out) return DEM_ERR_ARG;
  *out = nullptr;
  dem_engine *e = new dem_engine();
  *out = e;  // returned even on failure so that dem_last_error works; caller destroys it
  dem_engine *E = e;
  try {
    int ndev = 0;
    cudaError_t rc = cudaGetDeviceCount(&ndev);
    if (rc != cudaSuccess || ndev == 0) dem_fail(e, DEM_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU path", cudaGetErrorString(rc));
    if (device < 0 || device >= ndev) dem_fail(e, DEM_ERR_ARG, "device ordinal %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) dem_fail(e, DEM_ERR_CUDA, "device %d is sm_%d%d; this library only contains sm_100a code", device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    e->device = device; e->rank = rank; e->nranks = nranks; e->stream = (cudaStream_t)stream;
    if (nranks != 1) dem_fail(e, DEM_ERR_UNSUPPORTED, "multi-rank bricks are not enabled in this build yet");
    (void)nccl_id;
    CK(cudaHostAlloc((void **)&e->hflag, 4 * sizeof(int), cudaHostAllocMapped));
    e->hflag[0] = e->hflag[1] = 0;
    e->pm.tdamp = 1;
  } catch (const DemFail &f) { return f.code; }
  return DEM_OK;
}

extern "C" void dem_destroy(dem_engine *e)
{
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (int b = 0; b < 2; b++) { e->xr[b].release(); e->vm[b].release(); e->wt[b].release(); }
  e->xh.release(); e->tag.release(); e->tag_tmp.release(); e->density.release(); e->density_tmp.release();
  e->f.release(); e->tq.release(); e->whist.release(); e->whist_tmp.release(); e->tab.release(); e->dwalls.release();
  e->valid_tmp.release(); e->wlist.release(); e->fw.release(); e->gsrc.release(); e->gshift.release(); e->gsrc2.release(); e->gshift2.release();
  e->flo.release(); e->fhi.release(); e->slo.release(); e->shi.release(); e->ocs.release(); e->oce.release();
  e->gcs.release(); e->gce.release(); e->perm.release(); e->vals.release(); e->keys.release(); e->keys2.release();
  e->cubtmp.release(); e->overflow.release(); e->counters.release();
  for (int s = 0; s < 2; s++) { e->ls[s].nbr.release(); e->ls[s].ptag.release(); e->ls[s].numneigh.release(); e->ls[s].hist.release(); }
  for (auto &ev : e->ev) cudaEventDestroy(ev);
  if (e->hflag) cudaFreeHost(e->hflag);
  delete e;
}

extern "C" int dem_set_option(dem_engine *e, const char *name, double value)
{
  API_BEGIN
  e->opt[name] = value;
  API_END
}

extern "C" int dem_set_units(dem_engine *e, const char *s)
{
  API_BEGIN
  if (!strcmp(s, "si") || !strcmp(s, "cgs") || !strcmp(s, "micro")) { e->nktv2p = e->ftm2v = 1.0; }
  else dem_fail(e, DEM_ERR_UNSUPPORTED, "units %s not supported (si, cgs, micro)", s);
  API_END
}
extern "C" int dem_set_box(dem_engine *e, const double lo[3], const double hi[3], const int periodic[3])
{
  API_BEGIN
  for (int d = 0; d < 3; d++) {
    if (!(hi[d] > lo[d])) dem_fail(e, DEM_ERR_ARG, "box hi <= lo in dim %d", d);
    e->lo[d] = lo[d]; e->hi[d] = hi[d]; e->prd[d] = hi[d] - lo[d]; e->periodic[d] = periodic[d] ? 1 : 0;
  }
  API_END
}
extern "C" int dem_set_ntypes(dem_engine *e, int n)
{
  API_BEGIN
  if (n < 1 || n > MAXT) dem_fail(e, DEM_ERR_ARG, "ntypes must be in 1..%d", MAXT);
  e->ntypes = n;
  API_END
}
extern "C" int dem_set_processors(dem_engine *e, int px, int py, int pz)
{
  API_BEGIN
  if (px * py * pz != e->nranks) dem_fail(e, DEM_ERR_ARG, "processors grid %dx%dx%d != nranks %d", px, py, pz, e->nranks);
  API_END
}
extern "C" int dem_set_neighbor(dem_engine *e, double skin, int every, int delay, int check)
{
  API_BEGIN
  if (skin < 0 || every < 1 || delay < 0) dem_fail(e, DEM_ERR_ARG, "bad neighbor settings");
  e->skin = skin; e->every = every; e->delay = delay; e->check = check;
  API_END
}
extern "C" int dem_set_timestep(dem_engine *e, double dt)
{
  API_BEGIN
  if (!(dt > 0)) dem_fail(e, DEM_ERR_ARG, "timestep must be > 0");
  e->dt = dt;
  API_END
}

extern "C" int dem_set_property(dem_engine *e, const char *name, const char *kind, const double *v, int n)
{
  API_BEGIN
  const int T = e->ntypes;
  std::string nm(name), kd(kind);
  if (kd == "scalar") {
    if (nm == "characteristicVelocity" && n == 1) e->charVel = v[0];
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "scalar property %s not on the hot path", name);
  } else if (kd == "peratomtype") {
    if (n != T) dem_fail(e, DEM_ERR_ARG, "%s: peratomtype needs %d values", name, T);
    double *dst = nm == "youngsModulus" ? e->Y : nm == "poissonsRatio" ? e->nu : nullptr;
    if (!dst) dem_fail(e, DEM_ERR_UNSUPPORTED, "peratomtype property %s not on the hot path", name);
    for (int i = 0; i < T; i++) dst[i + 1] = v[i];
  } else if (kd == "peratomtypepair") {
    if (n != T * T) dem_fail(e, DEM_ERR_ARG, "%s: peratomtypepair needs %d values", name, T * T);
    double(*dst)[MAXT + 1] = nm == "coefficientRestitution" ? e->cor : nm == "coefficientFriction" ? e->mu
                            : nm == "coefficientRollingFriction" ? e->rmu : nm == "coefficientRollingViscousDamping" ? e->rvisc : nullptr;
    if (!dst) dem_fail(e, DEM_ERR_UNSUPPORTED, "peratomtypepair property %s not on the hot path", name);
    for (int i = 0; i < T; i++) for (int j = 0; j < T; j++) {
      if (v[i * T + j] != v[j * T + i]) dem_fail(e, DEM_ERR_ARG, "%s: per-atomtype property matrix must be symmetric", name);
      dst[i + 1][j + 1] = v[i * T + j];
    }
  } else dem_fail(e, DEM_ERR_ARG, "unknown property kind %s", kind);
  e->have_prop[nm] = 1;
  API_END
}

// model selection in the reference's fixed keyword order (contact_models.cpp:158-260)
static void parse_model_select(dem_engine *e, int &argc, const char *const *&a, ModelP &m)
{
  memset(&m, 0, sizeof m); m.tdamp = 1; m.off_shear = m.off_roll = -1;
  if (argc > 1 && !strcmp(a[0], "model")) {
    if (!strcmp(a[1], "hertz")) m.normal = N_HERTZ;
    else if (!strcmp(a[1], "hooke")) m.normal = N_HOOKE;
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "normal model '%s' is outside the hot-path scope (hertz, hooke)", a[1]);
    a += 2; argc -= 2;
  } else dem_fail(e, DEM_ERR_ARG, "expected 'model <normal model>'");
  if (argc > 1 && !strcmp(a[0], "tangential")) {
    if (!strcmp(a[1], "history")) m.tangential = 1;
    else if (!strcmp(a[1], "off")) m.tangential = 0;
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "tangential model '%s' is outside the hot-path scope (history)", a[1]);
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "cohesion")) {
    if (strcmp(a[1], "off")) dem_fail(e, DEM_ERR_UNSUPPORTED, "cohesion model '%s' not built yet", a[1]);
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "rolling_friction")) {
    if (!strcmp(a[1], "cdt")) m.rolling = R_CDT; else if (!strcmp(a[1], "epsd")) m.rolling = R_EPSD;
    else if (!strcmp(a[1], "epsd2")) m.rolling = R_EPSD2; else if (!strcmp(a[1], "off")) m.rolling = R_OFF;
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "rolling model '%s' is outside the hot-path scope (cdt, epsd, epsd2)", a[1]);
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "surface")) {
    if (strcmp(a[1], "default")) dem_fail(e, DEM_ERR_UNSUPPORTED, "surface model '%s' is outside the hot-path scope", a[1]);
    a += 2; argc -= 2;
  }
  if ((m.rolling == R_EPSD || m.rolling == R_EPSD2) && !m.tangential)
    dem_fail(e, DEM_ERR_ARG, "rolling_friction epsd/epsd2 requires tangential history");
  m.dnum = 0; m.hrec = 0; m.rec_shear = m.rec_roll = -1;
  if (m.tangential) { m.off_shear = m.dnum; m.dnum += 3; m.rec_shear = m.hrec++; }
  if (m.rolling == R_EPSD || m.rolling == R_EPSD2) { m.off_roll = m.dnum; m.dnum += 3; m.rec_roll = m.hrec++; }
}
// trailing `key on|off` settings (Settings::parseArguments)
static void parse_model_settings(dem_engine *e, int argc, const char *const *a, ModelP &m)
{
  while (argc > 0) {
    if (argc < 2) dem_fail(e, DEM_ERR_ARG, "Unknown argument or wrong keyword order: '%s'", a[0]);
    int on;
    if (!strcmp(a[1], "on")) on = 1; else if (!strcmp(a[1], "off")) on = 0;
    else { dem_fail(e, DEM_ERR_ARG, "Unknown argument or wrong keyword order: '%s'", a[0]); return; }
    if (!strcmp(a[0], "tangential_damping")) m.tdamp = on;
    else if (!strcmp(a[0], "limitForce")) m.limitForce = on;
    else if (!strcmp(a[0], "torsionTorque") && m.rolling != R_OFF) m.torsion = on;
    else if (!strcmp(a[0], "ktToKnUser") && m.normal == N_HOOKE) m.ktToKn = on;
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "setting '%s' is unknown or outside the hot-path scope", a[0]);
    a += 2; argc -= 2;
  }
}

extern "C" int dem_set_pair_style(dem_engine *e, int argc, const char *const *argv)
{
  API_BEGIN
  if (e->setup_done) dem_fail(e, DEM_ERR_STATE, "pair_style cannot change after setup");
  ModelP m;
  parse_model_select(e, argc, argv, m);
  parse_model_settings(e, argc, argv, m);
  e->pm = m; e->have_pair = 1;
  API_END
}

extern "C" int dem_add_wall_primitive(dem_engine *e, const char *id, int argc, const char *const *argv)
{
  API_BEGIN
  if (e->setup_done) dem_fail(e, DEM_ERR_STATE, "walls cannot be added after setup");
  if ((int)e->walls.size() == DEM_MAXW) dem_fail(e, DEM_ERR_OVERFLOW, "at most %d primitive walls", DEM_MAXW);
  for (auto &w : e->walls) if (w.id == id) dem_fail(e, DEM_ERR_ARG, "fix id %s already in use", id);
  WallHost W; W.id = id; memset(&W.p, 0, sizeof W.p);
  parse_model_select(e, argc, argv, W.p.m);
  if (argc < 4 || strcmp(argv[0], "primitive")) {
    if (argc > 0 && !strcmp(argv[0], "mesh")) dem_fail(e, DEM_ERR_UNSUPPORTED, "mesh walls go through dem_add_wall_mesh");
    dem_fail(e, DEM_ERR_ARG, "Need to use define style 'mesh' or 'primitive'");
  }
  if (strcmp(argv[1], "type")) dem_fail(e, DEM_ERR_ARG, "expecting keyword 'type'");
  W.p.atom_type = atoi(argv[2]);
  if (W.p.atom_type < 1 || W.p.atom_type > e->ntypes) dem_fail(e, DEM_ERR_ARG, "1 <= type <= max type as defined in create_box");
  static const char *names[6] = {"xplane", "yplane", "zplane", "xcylinder", "ycylinder", "zcylinder"};
  W.p.wtype = -1;
  for (int k = 0; k < 6; k++) if (!strcmp(argv[3], names[k])) W.p.wtype = k;
  if (W.p.wtype < 0) dem_fail(e, DEM_ERR_ARG, "unknown primitive wall style");
  const int np = W.p.wtype < 3 ? 1 : 3;
  if (argc < 4 + np) dem_fail(e, DEM_ERR_ARG, "not enough arguments for primitive wall");
  for (int k = 0; k < np; k++) W.p.param[k] = atof(argv[4 + k]);
  argv += 4 + np; argc -= 4 + np;
  W.p.shearAxis = -1;
  while (argc > 0) {
    if (!strcmp(argv[0], "shear")) {
      if (argc < 3) dem_fail(e, DEM_ERR_ARG, "not enough arguments for 'shear'");
      if (strlen(argv[1]) != 1 || argv[1][0] < 'x' || argv[1][0] > 'z') dem_fail(e, DEM_ERR_ARG, "illegal 'shear' dim");
      W.p.shearDim = argv[1][0] - 'x'; W.p.vshear = atof(argv[2]); W.p.shear = 1;
      const int axis = W.p.wtype >= 3 ? W.p.wtype - 3 : -1;
      if (W.p.shearDim != axis) { W.p.shearAxis = axis; if (axis >= 0) W.p.axisVec[axis] = W.p.vshear; }
      argv += 3; argc -= 3;
    } else if (!strcmp(argv[0], "temperature") || !strcmp(argv[0], "store_force") || !strcmp(argv[0], "store_force_contact"))
      dem_fail(e, DEM_ERR_UNSUPPORTED, "wall keyword '%s' is outside the hot-path scope", argv[0]);
    else break;
  }
  parse_model_settings(e, argc, argv, W.p.m);
  W.p.hist_row = e->nwrows;
  e->nwrows += W.p.m.dnum;
  e->walls.push_back(W);
  API_END
}

extern "C" int dem_set_gravity(dem_engine *e, double mag, const double dir[3])
{
  API_BEGIN
  const double len = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  if (len == 0.) dem_fail(e, DEM_ERR_ARG, "Gravity direction vector = 0");
  for (int d = 0; d < 3; d++) { const double u = dir[d] / len; e->g[d] = mag * u; }
  e->have_g = 1;
  API_END
}
extern "C" int dem_set_freeze(dem_engine *e, int bit) { API_BEGIN e->freezebit = bit; API_END }
extern "C" int dem_set_integrate(dem_engine *e, int bit) { API_BEGIN e->integbit = bit; API_END }

// ------------------------------------------------------------------------------------------------
static void ensure_particle_cap(dem_engine *E, long need, long keep)
{
  if (need <= E->cap) return;
  const int oldcap = E->cap;
  long ncap = std::max(need + need / 8 + 256, (long)oldcap * 3 / 2);
  ncap = (ncap + 127) / 128 * 128;
  cudaStream_t st = E->stream;
  for (int b = 0; b < 2; b++) { E->xr[b].ensure(E, ncap, keep, st); E->vm[b].ensure(E, ncap, keep, st); E->wt[b].ensure(E, ncap, keep, st); }
  E->xh.ensure(E, ncap, keep, st);
  E->tag.ensure(E, ncap, keep, st); E->tag_tmp.ensure(E, ncap, 0, st);
  E->density.ensure(E, ncap, keep, st); E->density_tmp.ensure(E, ncap, 0, st);
  E->valid_tmp.ensure(E, ncap, 0, st);
  // row-major [rows][cap] arrays: re-stride
  auto restride = [&](DevBuf<double> &b, int rows) {
    if (!rows) return;
    DevBuf<double> nb; nb.ensure(E, (size_t)rows * ncap, 0, st);
    CK(cudaMemsetAsync(nb.p, 0, (size_t)rows * ncap * sizeof(double), st));
    if (b.p && oldcap && keep)
      CK(cudaMemcpy2DAsync(nb.p, ncap * sizeof(double), b.p, (size_t)oldcap * sizeof(double), std::min<long>(keep, oldcap) * sizeof(double), rows, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    b.release(); b = nb;
  };
  restride(E->f, 3); restride(E->tq, 3); restride(E->whist, E->nwrows);
  if (E->nwrows) { E->whist_tmp.release(); E->whist_tmp.ensure(E, (size_t)E->nwrows * ncap, 0, st); }
  E->cap = (int)ncap;
}

extern "C" int dem_upload_particles(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x,
                                    const double *v, const double *omega, const double *radius, const double *density)
{
  API_BEGIN
  if (n < 0 || (n > 0 && (!tag || !type || !x || !radius || !density))) dem_fail(e, DEM_ERR_ARG, "missing particle arrays");
  if (n >= (long)NBR_IDX) dem_fail(e, DEM_ERR_OVERFLOW, "more than 2^30 particles on one GPU");
  CK(cudaSetDevice(e->device));
  e->cap = 0;  // force fresh allocation
  for (int b = 0; b < 2; b++) { e->xr[b].release(); e->vm[b].release(); e->wt[b].release(); }
  e->xh.release(); e->tag.release(); e->density.release(); e->f.release(); e->tq.release(); e->whist.release();
  ensure_particle_cap(e, std::max<long>(n + n / 4 + 1024, 1024), 0);
  std::vector<double4> hx(n), hv(n), hw(n);
  double rmax = 0.0;
  for (long i = 0; i < n; i++) {
    if (type[i] < 1 || type[i] > e->ntypes) dem_fail(e, DEM_ERR_ARG, "Invalid atom type in particle %ld", i);
    if (!(radius[i] > 0.0) || !(density[i] > 0.0)) dem_fail(e, DEM_ERR_ARG, "Invalid radius or density in particle %ld", i);
    if (tag[i] <= 0) dem_fail(e, DEM_ERR_ARG, "Invalid atom ID in particle %ld", i);
    const double r = radius[i];
    const double m = 4.0 * 3.14159265358979323846 / 3.0 * r * r * r * density[i];  // atom_vec_sphere.cpp:1078
    hx[i] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], r);
    hv[i] = make_double4(v ? v[3 * i] : 0., v ? v[3 * i + 1] : 0., v ? v[3 * i + 2] : 0., m);
    const long long bits = pack_bits(type[i], mask ? mask[i] : 1);
    double wb; memcpy(&wb, &bits, 8);
    hw[i] = make_double4(omega ? omega[3 * i] : 0., omega ? omega[3 * i + 1] : 0., omega ? omega[3 * i + 2] : 0., wb);
    rmax = std::max(rmax, r);
  }
  e->cur = 0;
  if (n) {
    CK(cudaMemcpyAsync(e->xr[0].p, hx.data(), n * sizeof(double4), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->vm[0].p, hv.data(), n * sizeof(double4), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->wt[0].p, hw.data(), n * sizeof(double4), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->tag.p, tag, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->density.p, density, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  }
  CK(cudaMemsetAsync(e->xh.p, 0, (size_t)e->cap * sizeof(double4), e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->nlocal = n; e->nghost = 0; e->rmax = rmax;
  e->uploaded = 1; e->setup_done = 0; e->forces_valid = 0;
  e->ls[0].valid = e->ls[1].valid = 0;
  e->ntimestep = 0;
  API_END
}

// ------------------------------------------------------------------------------------------------
static void derive_tables(dem_engine *E)
{  // global_properties.cpp:428-452 (Yeff), 458-483 (Geff), 519-537 (log e), 542-560 (betaeff)
  const int T = E->ntypes, n1 = T + 1;
  std::vector<double> t((size_t)T_COUNT * n1 * n1, 0.0);
  auto need = [&](const char *nm) { if (!E->have_prop.count(nm)) dem_fail(E, DEM_ERR_STATE, "property %s required by the selected models was not defined", nm); };
  bool hertz = false, hooke = false, tang = false, roll = false, epsd = false;
  auto scan = [&](const ModelP &m) { hertz |= m.normal == N_HERTZ; hooke |= m.normal == N_HOOKE; tang |= m.tangential != 0; roll |= m.rolling != R_OFF; epsd |= m.rolling == R_EPSD; };
  if (E->have_pair) scan(E->pm);
  for (auto &w : E->walls) scan(w.p.m);
  if (hertz || hooke) { need("youngsModulus"); need("poissonsRatio"); need("coefficientRestitution"); }
  if (hooke) need("characteristicVelocity");
  if (tang) need("coefficientFriction");
  if (roll) need("coefficientRollingFriction");
  if (epsd) need("coefficientRollingViscousDamping");
  for (int i = 1; i <= T; i++) for (int j = 1; j <= T; j++) {
    const double Yi = E->Y[i], Yj = E->Y[j], vi = E->nu[i], vj = E->nu[j];
    auto at = [&](int w) -> double & { return t[((size_t)w * n1 + i) * n1 + j]; };
    if (hertz || hooke) {
      at(T_YEFF) = 1. / ((1. - pow(vi, 2.)) / Yi + (1. - pow(vj, 2.)) / Yj);
      at(T_GEFF) = 1. / (2. * (2. - vi) * (1. + vi) / Yi + 2. * (2. - vj) * (1. + vj) / Yj);
      const double cr = E->cor[i][j];
      if (cr <= 0.05 || cr > 1) dem_fail(E, DEM_ERR_ARG, "0.05 < coefficientRestitution <= 1 required");
      at(T_CORLOG) = log(cr);
      at(T_BETA) = at(T_CORLOG) / sqrt(pow(at(T_CORLOG), 2.) + pow(3.14159265358979323846, 2.));
    }
    at(T_MU) = E->mu[i][j]; at(T_RMU) = E->rmu[i][j]; at(T_RVISC) = E->rvisc[i][j];
    if (hertz || hooke) { at(T_SQ2Y) = sqrt(2. * at(T_YEFF)); at(T_SQ8G) = sqrt(8. * at(T_GEFF)); at(T_INV8G) = 1. / (8. * at(T_GEFF)); }
  }
  for (int w = 0; w < T_COUNT; w++) E->t1[w] = t[((size_t)w * n1 + 1) * n1 + 1];
  E->tab.ensure(E, t.size());
  CK(cudaMemcpyAsync(E->tab.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, E->stream));
  if (!E->walls.empty()) {
    std::vector<WallP> hw;
    for (auto &w : E->walls) hw.push_back(w.p);
    E->dwalls.ensure(E, hw.size());
    CK(cudaMemcpyAsync(E->dwalls.p, hw.data(), hw.size() * sizeof(WallP), cudaMemcpyHostToDevice, E->stream));
  }
  CK(cudaStreamSynchronize(E->stream));
}

static void setup_grid(dem_engine *E)
{
  E->cutneighmax = 2.0 * E->rmax * E->cdf + E->skin;  // pair_gran.cpp:591-603 + neighbor skin
  if (!(E->cutneighmax > 0)) dem_fail(E, DEM_ERR_STATE, "neighbour cutoff is zero (no particles or zero radius)");
  long total = 1;
  double cell = E->cutneighmax;
  for (int pass = 0; pass < 64; pass++) {
    total = 1;
    for (int d = 0; d < 3; d++) {
      int nc = (int)floor(E->prd[d] / cell); if (nc < 1) nc = 1;
      const double size = E->prd[d] / nc;
      E->grid.nc[d] = nc + 2; E->grid.inv[d] = 1.0 / size; E->grid.org[d] = E->lo[d] - size;
      total *= (nc + 2);
    }
    if (total <= (1L << 25)) break;
    cell *= 1.3;
  }
  for (int d = 0; d < 3; d++)
    if (E->periodic[d] && E->prd[d] < 2.0 * E->cutneighmax)
      dem_fail(E, DEM_ERR_UNSUPPORTED, "periodic box length in dim %d is below two neighbour cutoffs", d);
  E->grid.morton = (E->grid.nc[0] <= 1024 && E->grid.nc[1] <= 1024 && E->grid.nc[2] <= 1024) ? 1 : 0;
  if (E->opt.count("morton") && E->opt["morton"] == 0) E->grid.morton = 0;
  E->ncells = total;
  E->ocs.ensure(E, total); E->oce.ensure(E, total); E->gcs.ensure(E, total); E->gce.ensure(E, total);
}

static void ensure_cub(dem_engine *E, size_t n)
{
  size_t b1 = 0, b2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (unsigned *)nullptr, (unsigned *)nullptr, (int *)nullptr, (int *)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, b2, (int *)nullptr, (int *)nullptr, (int)n);
  E->cubtmp.ensure(E, std::max(b1, b2) + 256);
}

static GhostP ghost_params(dem_engine *E)
{
  GhostP G;
  G.nghost = (int)E->nghost; G.nlocal = (int)E->nlocal; G.src = E->gsrc.p; G.shift = E->gshift.p;
  for (int d = 0; d < 3; d++) G.prd[d] = E->prd[d];
  G.xr = E->xr[E->cur].p; G.vm = E->vm[E->cur].p; G.wt = E->wt[E->cur].p; G.with_static = 1;
  return G;
}
static void ghost_update(dem_engine *E, int g0, int g1)
{
  if (g1 <= g0) return;
  GhostP G = ghost_params(E);
  k_ghost_update<<<GRID(g1 - g0, 256), 256, 0, E->stream>>>(G, g0, g1);
  E->launches++;
}

static void ensure_list(dem_engine *E, ListSet &L, int cap, int maxk, int dnum, int hslots)
{
  if (L.cap != cap || L.maxk < maxk || L.dnum != dnum || L.hslots < hslots) {
    L.nbr.release(); L.ptag.release(); L.hist.release(); L.numneigh.release();
    L.cap = cap; L.maxk = maxk; L.dnum = dnum; L.hslots = hslots;
    L.nbr.ensure(E, (size_t)maxk * cap); L.ptag.ensure(E, (size_t)maxk * cap); L.numneigh.ensure(E, cap);
    if (dnum) L.hist.ensure(E, (size_t)hslots * dnum * cap);
  }
}

// Neighbor rebuild: verlet.cpp:305-328 (pre_exchange .. neighbor->build) re-designed for the GPU.
static void rebuild(dem_engine *E)
{
  cudaStream_t st = E->stream;
  const int n = (int)E->nlocal;
  const int dnum = E->have_pair ? E->pm.hrec : 0;  // history records per contact
  ensure_cub(E, (size_t)E->cap);
  E->keys.ensure(E, E->cap); E->keys2.ensure(E, E->cap); E->vals.ensure(E, E->cap); E->perm.ensure(E, E->cap);
  E->overflow.ensure(E, 2);
  BoxP B;
  for (int d = 0; d < 3; d++) { B.lo[d] = E->lo[d]; B.hi[d] = E->hi[d]; B.prd[d] = E->prd[d]; B.periodic[d] = E->periodic[d]; }
  int c = E->cur;
  if (n) {
    // 1. pbc wrap, cell keys, radix sort, gather the particle records into cell (Morton) order
    k_wrap_key<<<GRID(n, 256), 256, 0, st>>>(n, E->xr[c].p, E->grid, B, E->keys.p, E->vals.p);
    size_t tb = E->cubtmp.n;
    int endbit = 32;
    if (E->grid.morton) endbit = 30; else { endbit = 1; while ((1L << endbit) < E->ncells) endbit++; }
    CK(cub::DeviceRadixSort::SortPairs(E->cubtmp.p, tb, E->keys.p, E->keys2.p, E->vals.p, E->perm.p, n, 0, endbit, st));
    k_gather4<<<GRID(n, 256), 256, 0, st>>>(n, E->perm.p, E->xr[c].p, E->xr[c ^ 1].p, E->vm[c].p, E->vm[c ^ 1].p, E->wt[c].p, E->wt[c ^ 1].p);
    k_gather_rows<int><<<GRID(n, 256), 256, 0, st>>>(n, 1, 0, 0, E->perm.p, E->tag.p, E->tag_tmp.p);
    k_gather_rows<double><<<GRID(n, 256), 256, 0, st>>>(n, 1, 0, 0, E->perm.p, E->density.p, E->density_tmp.p);
    std::swap(E->tag.p, E->tag_tmp.p); std::swap(E->tag.n, E->tag_tmp.n);
    std::swap(E->density.p, E->density_tmp.p); std::swap(E->density.n, E->density_tmp.n);
    if (E->nwrows) {
      k_gather_rows<double><<<GRID(n, 256), 256, 0, st>>>(n, E->nwrows, (size_t)E->cap, (size_t)E->cap, E->perm.p, E->whist.p, E->whist_tmp.p);
      std::swap(E->whist.p, E->whist_tmp.p); std::swap(E->whist.n, E->whist_tmp.n);
      E->launches++;
    }
    k_extract_valid<<<GRID(n, 256), 256, 0, st>>>(n, E->perm.p, E->xh.p, E->valid_tmp.p);
    E->launches += 6;
    E->cur = c ^ 1; c = E->cur;
  }
  // 2. owned cell ranges
  CK(cudaMemsetAsync(E->ocs.p, 0, E->ncells * sizeof(int), st)); CK(cudaMemsetAsync(E->oce.p, 0, E->ncells * sizeof(int), st));
  CK(cudaMemsetAsync(E->gcs.p, 0, E->ncells * sizeof(int), st)); CK(cudaMemsetAsync(E->gce.p, 0, E->ncells * sizeof(int), st));
  if (n) { k_cell_ranges<<<GRID(n, 256), 256, 0, st>>>(n, 0, E->xr[c].p, E->grid, E->ocs.p, E->oce.p); E->launches++; }
  // 3. periodic images (ghosts), one dimension after the other so that edges/corners propagate
  E->nghost = 0;
  for (int d = 0; d < 3 && n; d++) {
    if (!E->periodic[d]) continue;
    const int n0 = (int)(E->nlocal + E->nghost);
    E->flo.ensure(E, n0 + 1); E->fhi.ensure(E, n0 + 1); E->slo.ensure(E, n0 + 1); E->shi.ensure(E, n0 + 1);
    k_border_flag<<<GRID(n0, 256), 256, 0, st>>>(n0, E->xr[c].p, d, E->lo[d] + E->cutneighmax, E->hi[d] - E->cutneighmax, E->flo.p, E->fhi.p);
    size_t tb = E->cubtmp.n;
    CK(cub::DeviceScan::ExclusiveSum(E->cubtmp.p, tb, E->flo.p, E->slo.p, n0, st));
    tb = E->cubtmp.n;
    CK(cub::DeviceScan::ExclusiveSum(E->cubtmp.p, tb, E->fhi.p, E->shi.p, n0, st));
    int last[4];
    CK(cudaMemcpyAsync(&last[0], E->flo.p + n0 - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last[1], E->slo.p + n0 - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last[2], E->fhi.p + n0 - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&last[3], E->shi.p + n0 - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int nlo = last[0] + last[1], nhi = last[2] + last[3];
    E->launches += 3;
    if (nlo + nhi == 0) continue;
    ensure_particle_cap(E, (long)n0 + nlo + nhi, n0);
    const long ng = E->nghost + nlo + nhi;
    E->gsrc.ensure(E, ng + 1, E->nghost, st); E->gshift.ensure(E, 3 * (ng + 1), 3 * E->nghost, st);
    k_border_scatter<<<GRID(n0, 256), 256, 0, st>>>(n0, n, d, E->flo.p, E->slo.p, E->fhi.p, E->shi.p, nlo, (int)E->nghost, E->gsrc.p, E->gshift.p);
    E->launches++;
    const int g0 = (int)E->nghost;
    E->nghost = ng;
    ghost_update(E, g0, (int)ng);
  }
  // 4. ghosts into cell order
  if (E->nghost) {
    const int ng = (int)E->nghost;
    E->keys.ensure(E, E->cap); E->keys2.ensure(E, E->cap); E->vals.ensure(E, E->cap);
    DevBuf<int> &gperm = E->slo;  // reuse
    gperm.ensure(E, ng);
    ensure_cub(E, (size_t)E->cap);
    k_ghost_keys<<<GRID(ng, 256), 256, 0, st>>>(ng, n, E->xr[c].p, E->grid, E->keys.p, E->vals.p);
    size_t tb = E->cubtmp.n;
    int endbit = 32;
    if (E->grid.morton) endbit = 30; else { endbit = 1; while ((1L << endbit) < E->ncells) endbit++; }
    CK(cub::DeviceRadixSort::SortPairs(E->cubtmp.p, tb, E->keys.p, E->keys2.p, E->vals.p, gperm.p, ng, 0, endbit, st));
    E->gsrc2.ensure(E, ng); E->gshift2.ensure(E, 3 * (size_t)ng);
    E->tag.ensure(E, E->cap, E->nlocal, st);
    k_ghost_permute<<<GRID(ng, 256), 256, 0, st>>>(ng, gperm.p, E->gsrc.p, E->gshift.p, E->gsrc2.p, E->gshift2.p, E->tag.p, E->tag.p, n);
    std::swap(E->gsrc.p, E->gsrc2.p); std::swap(E->gsrc.n, E->gsrc2.n);
    std::swap(E->gshift.p, E->gshift2.p); std::swap(E->gshift.n, E->gshift2.n);
    E->launches += 3;
    ghost_update(E, 0, ng);
    k_cell_ranges<<<GRID(ng, 256), 256, 0, st>>>(ng, n, E->xr[c].p, E->grid, E->gcs.p, E->gce.p);
    E->launches++;
  }
  // the spare record buffers must be as large as the live ones (ghost growth may have re-allocated)
  // 5. full Verlet list + history remap
  ListSet &Lold = E->ls[E->lcur], &Lnew = E->ls[E->lcur ^ 1];
  int maxk = std::max(Lold.valid ? Lold.maxk : 0, (int)(E->opt.count("maxneigh") ? E->opt["maxneigh"] : 24));
  int hslots = std::max(Lold.valid ? Lold.hslots : 0, (int)(E->opt.count("histslots") ? E->opt["histslots"] : 16));
  for (int attempt = 0; attempt < 6 && n; attempt++) {
    ensure_list(E, Lnew, E->cap, maxk, dnum, hslots);
    CK(cudaMemsetAsync(E->overflow.p, 0, 2 * sizeof(int), st));
    BuildP P;
    P.nlocal = n; P.cap = Lnew.cap; P.maxk = Lnew.maxk; P.dnum = dnum; P.hslots = Lnew.hslots; P.xr = E->xr[c].p; P.tag = E->tag.p; P.G = E->grid;
    P.ocs = E->ocs.p; P.oce = E->oce.p; P.gcs = E->gcs.p; P.gce = E->gce.p; P.cdf = E->cdf; P.skin = E->skin;
    P.nbr = Lnew.nbr.p; P.numneigh = Lnew.numneigh.p; P.ptag = Lnew.ptag.p; P.hist = Lnew.hist.p;
    P.have_old = (Lold.valid && dnum && Lold.dnum == dnum) ? 1 : 0; P.cap_old = Lold.cap; P.dnum_old = Lold.dnum;
    P.perm = E->perm.p; P.nbr_old = Lold.nbr.p; P.numneigh_old = Lold.numneigh.p; P.ptag_old = Lold.ptag.p; P.hist_old = Lold.hist.p;
    P.overflow = E->overflow.p;
    k_build_list<<<GRID(n, 128), 128, 0, st>>>(P);
    E->launches++;
    int ov[2] = {0, 0};
    CK(cudaMemcpyAsync(ov, E->overflow.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ov[0] == 0 && ov[1] == 0) break;
    if (attempt == 5) dem_fail(E, DEM_ERR_OVERFLOW, "neighbour list overflow (%d neighbours, %d history partners)", ov[0], ov[1]);
    if (ov[0]) maxk = ov[0] + 4;
    if (ov[1]) hslots = ov[1] + 12;
    if (maxk > 0xffff || hslots > NBR_MAXSLOTS) dem_fail(E, DEM_ERR_OVERFLOW, "a particle has %d neighbours / %d history partners", ov[0], ov[1]);
  }
  Lnew.valid = 1; Lold.valid = 0;
  E->lcur ^= 1;
  // 6. positions at build time + primitive wall candidate bits
  if (n) {
    const int nw = (int)E->walls.size();
    E->flo.ensure(E, n + 1); E->slo.ensure(E, n + 1);
    k_hold<<<GRID(n, 256), 256, 0, st>>>(n, E->xr[c].p, E->xh.p, E->valid_tmp.p, E->dwalls.p, nw, E->skin, nw ? E->flo.p : nullptr);
    E->launches++;
    E->nwc = 0;
    if (nw) {
      size_t tb = E->cubtmp.n;
      CK(cub::DeviceScan::ExclusiveSum(E->cubtmp.p, tb, E->flo.p, E->slo.p, n, st));
      int last[2];
      CK(cudaMemcpyAsync(&last[0], E->flo.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(&last[1], E->slo.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      E->nwc = last[0] + last[1];
      if (E->nwc > E->nwcap) { E->nwcap = E->nwc + E->nwc / 4 + 256; E->wlist.release(); E->fw.release(); E->wlist.ensure(E, E->nwcap); E->fw.ensure(E, 6 * (size_t)E->nwcap); }
      if (E->nwc) k_wall_index<<<GRID(n, 256), 256, 0, st>>>(n, E->flo.p, E->slo.p, E->wlist.p, E->xh.p);
      E->launches += 2;
    }
  }
  CK(cudaGetLastError());
  E->ago = 0;
  E->nbuilds++;
}

static StepP step_params(dem_engine *E, int mode)
{
  StepP P;
  memset(&P, 0, sizeof P);
  const int c = E->cur;
  ListSet &L = E->ls[E->lcur];
  P.nlocal = (int)E->nlocal; P.nall = (int)(E->nlocal + E->nghost); P.cap = E->cap; P.maxk = L.maxk; P.lcap = L.cap;
  P.xr = E->xr[c].p; P.vm = E->vm[c].p; P.wt = E->wt[c].p;
  P.xr_o = E->xr[c ^ 1].p; P.vm_o = E->vm[c ^ 1].p; P.wt_o = E->wt[c ^ 1].p;
  P.xh = E->xh.p; P.nbr = L.nbr.p; P.numneigh = L.numneigh.p; P.hist = L.hist.p; P.hslots = L.hslots;
  P.whist = E->whist.p; P.f = E->f.p; P.tq = E->tq.p; P.walls = E->dwalls.p; P.nwalls = (int)E->walls.size(); P.nwc = E->nwc; P.nwcap = E->nwcap; P.wlist = E->wlist.p; P.fw = E->fw.p;
  P.pm = E->pm; P.tab = E->tab.p; P.nt1 = E->ntypes + 1;
  for (int w = 0; w < T_COUNT; w++) P.t1[w] = E->t1[w];
  P.dt = E->dt; P.dtv = E->dt; P.dtf = 0.5 * E->dt * E->ftm2v; P.dtfrot = P.dtf / 0.4;  // fix_nve.cpp:86, fix_nve_sphere.cpp:69,150
  P.nktv2p = E->nktv2p; P.charVel = E->charVel; P.cdf = E->cdf; P.cdfsq = E->cdf * E->cdf;
  P.trigsq = 0.25 * E->skin * E->skin;  // neighbor.cpp:298
  P.cutneighmax = E->cutneighmax;
  for (int d = 0; d < 3; d++) P.g[d] = E->g[d];
  P.have_g = E->have_g; P.have_pair = E->have_pair; P.freezebit = E->freezebit; P.integbit = E->integbit;
  P.mode = mode; P.debug = E->opt.count("debug") ? (int)E->opt["debug"] : 0; P.flag = E->hflag; P.ncontact = nullptr;
  return P;
}

template <int N, int R>
static void launch_step_t(dem_engine *E, const StepP &P)
{
  if (E->ntypes == 1) k_step<N, R, true><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
  else k_step<N, R, false><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
}
static void launch_step(dem_engine *E, int mode, bool timed)
{
  if (!E->nlocal) return;
  StepP P = step_params(E, mode);
  const bool tm = timed && E->opt.count("time_kernels") && E->opt["time_kernels"] != 0;
  if (tm) {
    if ((long)E->ev.size() < 2 * (E->ev_used + 1)) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); E->ev.push_back(a); E->ev.push_back(b); }
    cudaEventRecord(E->ev[2 * E->ev_used], E->stream);
  }
  if (P.nwc) { k_walls<<<GRID(P.nwc, 128), 128, 0, E->stream>>>(P); E->launches++; }
  const int key = (E->have_pair ? E->pm.normal : N_HERTZ) * 4 + (E->have_pair ? E->pm.rolling : R_OFF);
  switch (key) {
    case N_HERTZ * 4 + R_OFF: launch_step_t<N_HERTZ, R_OFF>(E, P); break;
    case N_HERTZ * 4 + R_CDT: launch_step_t<N_HERTZ, R_CDT>(E, P); break;
    case N_HERTZ * 4 + R_EPSD: launch_step_t<N_HERTZ, R_EPSD>(E, P); break;
    case N_HERTZ * 4 + R_EPSD2: launch_step_t<N_HERTZ, R_EPSD2>(E, P); break;
    case N_HOOKE * 4 + R_OFF: launch_step_t<N_HOOKE, R_OFF>(E, P); break;
    case N_HOOKE * 4 + R_CDT: launch_step_t<N_HOOKE, R_CDT>(E, P); break;
    case N_HOOKE * 4 + R_EPSD: launch_step_t<N_HOOKE, R_EPSD>(E, P); break;
    case N_HOOKE * 4 + R_EPSD2: launch_step_t<N_HOOKE, R_EPSD2>(E, P); break;
    default: dem_fail(E, DEM_ERR_STATE, "no kernel for this model combination");
  }
  if (tm) { cudaEventRecord(E->ev[2 * E->ev_used + 1], E->stream); E->ev_used++; }
  E->launches++;
}

static void collect_timing(dem_engine *E)
{
  for (long k = 0; k < E->ev_used; k++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, E->ev[2 * k], E->ev[2 * k + 1]) == cudaSuccess) { E->step_ms += ms; E->step_calls++; }
  }
  E->ev_used = 0;
}

extern "C" int dem_setup(dem_engine *e)
{
  API_BEGIN
  if (!e->uploaded) dem_fail(e, DEM_ERR_STATE, "setup before dem_upload_particles");
  if (!(e->dt > 0)) dem_fail(e, DEM_ERR_STATE, "timestep not set");
  if (!e->have_pair && e->walls.empty()) dem_fail(e, DEM_ERR_STATE, "no pair_style and no wall defined");
  CK(cudaSetDevice(e->device));
  if (!e->setup_done) {
    derive_tables(e);
    setup_grid(e);
    if (e->nwrows) {
      e->whist.release(); e->whist.ensure(e, (size_t)e->nwrows * e->cap);
      CK(cudaMemsetAsync(e->whist.p, 0, (size_t)e->nwrows * e->cap * sizeof(double), e->stream));
      e->whist_tmp.release(); e->whist_tmp.ensure(e, (size_t)e->nwrows * e->cap);
    }
  }
  rebuild(e);
  e->nbuilds = 0;  // neighbor->ncalls counts the builds of the current run only
  launch_step(e, MODE_SETUP, false);
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaGetLastError());
  e->setup_done = 1; e->forces_valid = 1;
  API_END
}

extern "C" int dem_run(dem_engine *e, long nsteps)
{
  API_BEGIN
  if (!e->setup_done) dem_fail(e, DEM_ERR_STATE, "dem_run before dem_setup");
  if (nsteps < 0) dem_fail(e, DEM_ERR_ARG, "nsteps < 0");
  if (nsteps == 0) return DEM_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  e->step_ms = 0; e->step_calls = 0; e->ev_used = 0;
  // first half step from the stored forces (fix_nve_sphere.cpp:134-183)
  *e->hflag = 0;
  if (e->nlocal) {
    StepP P = step_params(e, MODE_STEP);
    k_initial_integrate<<<GRID(P.nlocal, 256), 256, 0, st>>>(P);
    e->launches++;
    e->cur ^= 1;
    ghost_update(e, 0, (int)e->nghost);
  }
  for (long s = 1; s <= nsteps; s++) {
    e->ntimestep++;
    // Neighbor::decide (neighbor.cpp:1362-1376)
    e->ago++;
    int nflag = 0;
    if (e->ago >= e->delay && e->ago % e->every == 0) {
      if (!e->check) nflag = 1;
      else { CK(cudaStreamSynchronize(st)); nflag = *e->hflag; }
    }
    if (nflag) { rebuild(e); CK(cudaStreamSynchronize(st)); *e->hflag = 0; }
    launch_step(e, s == nsteps ? MODE_LAST : MODE_STEP, true);
    e->cur ^= 1;
    ghost_update(e, 0, (int)e->nghost);
    if (e->ev_used >= 2048) { CK(cudaStreamSynchronize(st)); collect_timing(e); }
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  collect_timing(e);
  if (e->hflag[1]) { e->hflag[1] = 0; dem_fail(e, DEM_ERR_OVERFLOW, "a particle gained more new contacts between two rebuilds than free history slots; raise option 'histslots'"); }
  e->forces_valid = 1;
  API_END
}

// ------------------------------------------------------------------------------------------------
extern "C" long dem_nlocal(const dem_engine *e) { return e ? e->nlocal : 0; }

static std::vector<int> tag_order(dem_engine *E, std::vector<int> &tags)
{
  const long n = E->nlocal;
  tags.resize(n);
  if (n) CK(cudaMemcpy(tags.data(), E->tag.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int> o(n);
  for (long i = 0; i < n; i++) o[i] = (int)i;
  std::sort(o.begin(), o.end(), [&](int a, int b) { return tags[a] < tags[b]; });
  return o;
}

extern "C" int dem_download(dem_engine *e, const char *field, void *out, long count)
{
  API_BEGIN
  if (!e->uploaded) dem_fail(e, DEM_ERR_STATE, "no particles");
  if (count != e->nlocal) dem_fail(e, DEM_ERR_ARG, "count %ld != nlocal %ld", count, e->nlocal);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const long n = e->nlocal;
  std::vector<int> tags; std::vector<int> o = tag_order(e, tags);
  std::string f(field);
  const int c = e->cur;
  auto rec = [&](DevBuf<double4> &b) { std::vector<double4> h(n); if (n) CK(cudaMemcpy(h.data(), b.p, n * sizeof(double4), cudaMemcpyDeviceToHost)); return h; };
  if (f == "tag") { for (long k = 0; k < n; k++) ((int *)out)[k] = tags[o[k]]; }
  else if (f == "type" || f == "mask") {
    auto h = rec(e->wt[c]);
    for (long k = 0; k < n; k++) { long long b; memcpy(&b, &h[o[k]].w, 8); ((int *)out)[k] = f == "type" ? (int)(b & 0xff) : (int)((b >> 8) & 0xffffffffLL); }
  } else if (f == "radius") { auto h = rec(e->xr[c]); for (long k = 0; k < n; k++) ((double *)out)[k] = h[o[k]].w; }
  else if (f == "rmass") { auto h = rec(e->vm[c]); for (long k = 0; k < n; k++) ((double *)out)[k] = h[o[k]].w; }
  else if (f == "density") { std::vector<double> h(n); if (n) CK(cudaMemcpy(h.data(), e->density.p, n * sizeof(double), cudaMemcpyDeviceToHost)); for (long k = 0; k < n; k++) ((double *)out)[k] = h[o[k]]; }
  else if (f == "x" || f == "v" || f == "omega") {
    auto h = rec(f == "x" ? e->xr[c] : f == "v" ? e->vm[c] : e->wt[c]);
    for (long k = 0; k < n; k++) { ((double *)out)[3 * k] = h[o[k]].x; ((double *)out)[3 * k + 1] = h[o[k]].y; ((double *)out)[3 * k + 2] = h[o[k]].z; }
  } else if (f == "f" || f == "torque") {
    if (!e->forces_valid) dem_fail(e, DEM_ERR_STATE, "forces are only available after setup or run");
    std::vector<double> h(3 * (size_t)e->cap);
    CK(cudaMemcpy(h.data(), f == "f" ? e->f.p : e->tq.p, 3 * (size_t)e->cap * sizeof(double), cudaMemcpyDeviceToHost));
    for (long k = 0; k < n; k++) for (int d = 0; d < 3; d++) ((double *)out)[3 * k + d] = h[(size_t)d * e->cap + o[k]];
  } else dem_fail(e, DEM_ERR_ARG, "unknown field %s", field);
  API_END
}

struct PairRow { int lo, hi, flag; long src; };
static void collect_pairs(dem_engine *E, std::vector<PairRow> &rows, std::vector<double4> &hist, int &dnum)
{
  ListSet &L = E->ls[E->lcur];
  dnum = L.dnum;
  rows.clear();
  if (!L.valid || !E->nlocal) return;
  CK(cudaStreamSynchronize(E->stream));
  const long n = E->nlocal;
  std::vector<int> tags(n), nn(n);
  CK(cudaMemcpy(tags.data(), E->tag.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nn.data(), L.numneigh.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  for (long i = 0; i < n; i++) nn[i] &= 0xffff;
  int kmax = 0; for (long i = 0; i < n; i++) kmax = std::max(kmax, nn[i]);
  std::vector<unsigned> nbr((size_t)kmax * L.cap); std::vector<int> ptag((size_t)kmax * L.cap);
  if (kmax) {
    CK(cudaMemcpy(nbr.data(), L.nbr.p, nbr.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ptag.data(), L.ptag.p, ptag.size() * sizeof(int), cudaMemcpyDeviceToHost));
  }
  hist.assign((size_t)L.hslots * dnum * L.cap, make_double4(0., 0., 0., 0.));
  if (kmax && dnum) CK(cudaMemcpy(hist.data(), L.hist.p, hist.size() * sizeof(double4), cudaMemcpyDeviceToHost));
  for (long i = 0; i < n; i++) for (int k = 0; k < nn[i]; k++) {
    const unsigned w = nbr[(size_t)k * L.cap + i];
    const int tj = ptag[(size_t)k * L.cap + i];
    const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
    if (tags[i] < tj) rows.push_back(PairRow{tags[i], tj, slot >= 0 ? 1 : 0, (long)std::max(slot, 0) * L.cap + i});
  }
  std::sort(rows.begin(), rows.end(), [](const PairRow &a, const PairRow &b) { return a.lo != b.lo ? a.lo < b.lo : a.hi < b.hi; });
}

extern "C" int dem_pair_count(dem_engine *e, long *npairs, int *dnum)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  std::vector<PairRow> rows; std::vector<double4> hist; int dn = 0;
  collect_pairs(e, rows, hist, dn);
  if (npairs) *npairs = (long)rows.size();
  if (dnum) *dnum = e->have_pair ? e->pm.dnum : 0;
  API_END
}
extern "C" int dem_download_pairs(dem_engine *e, int *lo, int *hi, int *flag, double *hist)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  std::vector<PairRow> rows; std::vector<double4> h; int nrec = 0;
  collect_pairs(e, rows, h, nrec);
  ListSet &L = e->ls[e->lcur];
  const ModelP &M = e->pm;
  const int dn = e->have_pair ? M.dnum : 0;
  for (size_t r = 0; r < rows.size(); r++) {
    if (lo) lo[r] = rows[r].lo;
    if (hi) hi[r] = rows[r].hi;
    if (flag) flag[r] = rows[r].flag;
    if (hist) {
      const long k = rows[r].src / L.cap, i = rows[r].src % L.cap;  // k = history slot
      for (int d = 0; d < dn; d++) hist[r * dn + d] = 0.0;
      if (rows[r].flag) {
        if (M.rec_shear >= 0) { const double4 v = h[(size_t)(k * nrec + M.rec_shear) * L.cap + i]; hist[r * dn + M.off_shear] = v.x; hist[r * dn + M.off_shear + 1] = v.y; hist[r * dn + M.off_shear + 2] = v.z; }
        if (M.rec_roll >= 0) { const double4 v = h[(size_t)(k * nrec + M.rec_roll) * L.cap + i]; hist[r * dn + M.off_roll] = v.x; hist[r * dn + M.off_roll + 1] = v.y; hist[r * dn + M.off_roll + 2] = v.z; }
      }
    }
  }
  API_END
}

extern "C" int dem_download_wall_history(dem_engine *e, const char *wall_id, double *out, long count)
{
  API_BEGIN
  if (count != e->nlocal) dem_fail(e, DEM_ERR_ARG, "count %ld != nlocal %ld", count, e->nlocal);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  int wi = -1;
  for (size_t w = 0; w < e->walls.size(); w++) if (e->walls[w].id == wall_id) wi = (int)w;
  if (wi < 0) dem_fail(e, DEM_ERR_ARG, "no primitive wall with id %s", wall_id);
  const WallP &W = e->walls[wi].p;
  const long n = e->nlocal;
  std::vector<int> tags; std::vector<int> o = tag_order(e, tags);
  std::vector<double4> xh(n);
  if (n) CK(cudaMemcpy(xh.data(), e->xh.p, n * sizeof(double4), cudaMemcpyDeviceToHost));
  std::vector<double> h((size_t)W.m.dnum * e->cap);
  if (W.m.dnum && e->whist.p) CK(cudaMemcpy(h.data(), e->whist.p + (size_t)W.hist_row * e->cap, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (long k = 0; k < n; k++) {
    long long b; memcpy(&b, &xh[o[k]].w, 8);
    const bool valid = ((unsigned)(b & 0xffffffffLL) >> (16 + wi)) & 1u;
    for (int d = 0; d < W.m.dnum; d++) out[k * W.m.dnum + d] = valid ? h[(size_t)d * e->cap + o[k]] : 0.0;
  }
  API_END
}

extern "C" int dem_get_stats(dem_engine *e, dem_stats *s)
{
  API_BEGIN
  if (!s) dem_fail(e, DEM_ERR_ARG, "null stats");
  CK(cudaSetDevice(e->device));
  memset(s, 0, sizeof *s);
  s->ntimestep = e->ntimestep; s->nbuilds = e->nbuilds; s->nlocal = e->nlocal; s->nghost = e->nghost;
  s->kernel_launches = e->launches;
  ListSet &L = e->ls[e->lcur];
  s->maxneigh = L.maxk; s->dnum = e->have_pair ? e->pm.dnum : 0;
  if (L.valid && e->nlocal) {
    e->counters.ensure(e, 2);
    CK(cudaMemsetAsync(e->counters.p, 0, 2 * sizeof(unsigned long long), e->stream));
    k_count_pairs<<<GRID(e->nlocal, 256), 256, 0, e->stream>>>((int)e->nlocal, L.numneigh.p, L.nbr.p, L.cap, e->counters.p);
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, e->counters.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    s->npairs_full = (long)h[0]; s->ncontacts_full = (long)h[1];
  }
  s->step_kernel_ms = e->step_ms; s->step_kernel_calls = e->step_calls;
  API_END
}
