// dem_engine.cu -- host orchestration + C ABI (include/dem_b200.h) of the B200 DEM engine.
// One dem_engine == one GPU.  The step loop mirrors Verlet::run (verlet.cpp:264-391) but the
// per-step work is ONE fused kernel (dem_kernels.cuh) plus a small ghost refresh; rebuilds
// run the sort / border / list kernels.  No CPU fallback exists anywhere in this file.
#include <cub/cub.cuh>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "../../include/dem_b200.h"
#include "dem_pairs.cuh"
#include "dem_mesh_host.h"

using namespace dem;

#define MAXT 8

struct DemFail { int code; };

// NCCL is bound at run time (dlopen) so that single-GPU use needs no NCCL at all and multi-GPU use
// shares the copy torch.distributed already loaded into the process.
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load()
  {
    if (h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int k = 0; names[k] && !h; k++) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define NCCL_SYM(f) f = (decltype(f))dlsym(h, "nccl" #f); if (!f) return false;
    NCCL_SYM(GetUniqueId) NCCL_SYM(CommInitRank) NCCL_SYM(CommDestroy) NCCL_SYM(Send) NCCL_SYM(Recv) NCCL_SYM(AllReduce) NCCL_SYM(AllGather)
    NCCL_SYM(GroupStart) NCCL_SYM(GroupEnd) NCCL_SYM(GetErrorString)
#undef NCCL_SYM
    return true;
  }
};
static NcclApi g_nccl;
static void dem_fail(dem_engine *e, int code, const char *fmt, ...);

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess) { if (E) E->comm_bad = 1; dem_fail(E, DEM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); } \
  } while (0)

// Device memory comes from a process-wide cache of whole cudaMalloc blocks (per device, best fit within 25 %): engines
// that follow one another in a process (a parameter sweep, bench.py's two passes) reuse the blocks instead of paying the
// driver's map/unmap cost again -- cudaMalloc/cudaFree of the ~8 GB a 4M-sphere bed needs was measured at 0.2-0.6 s.
// Blocks are never sub-divided, so CUDA-IPC handles (peer-memory halo) stay valid.  dem_trim_memory() empties the cache.
static std::mutex g_mem_mu;
static std::multimap<size_t, void *> g_mem_free[32];
static std::map<void *, std::pair<int, size_t>> g_mem_size;
static bool mem_cache_on() { static const bool on = !getenv("DEM_B200_NO_CACHE"); return on; }
static void ipc_close_all();
static size_t mem_trim(int dev)
{
  size_t bytes = 0;
  ipc_close_all();
  std::lock_guard<std::mutex> lk(g_mem_mu);
  for (int d = 0; d < 32; d++) if (dev < 0 || d == dev) {
    int cur = 0; cudaGetDevice(&cur);
    if (!g_mem_free[d].empty()) cudaSetDevice(d);
    for (auto &kv : g_mem_free[d]) { bytes += kv.first; g_mem_size.erase(kv.second); cudaFree(kv.second); }
    if (!g_mem_free[d].empty()) cudaSetDevice(cur);
    g_mem_free[d].clear();
  }
  return bytes;
}
static cudaError_t dev_alloc(void **out, size_t bytes)
{
  int dev = 0; cudaGetDevice(&dev); dev &= 31;
  const size_t gran = bytes >= (1u << 20) ? (2u << 20) : 512;
  const size_t sz = (bytes + gran - 1) / gran * gran;
  if (mem_cache_on()) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_mem_free[dev].lower_bound(sz);
    if (it != g_mem_free[dev].end() && it->first <= sz + sz / 4 + (2u << 20)) { *out = it->second; g_mem_free[dev].erase(it); return cudaSuccess; }
  }
  cudaError_t rc = cudaMalloc(out, sz);
  if (rc != cudaSuccess) { cudaGetLastError(); mem_trim(dev); rc = cudaMalloc(out, sz); }
  if (rc == cudaSuccess) { std::lock_guard<std::mutex> lk(g_mem_mu); g_mem_size[*out] = {dev, sz}; }
  return rc;
}
static void dev_free(void *p)
{
  if (!p) return;
  cudaDeviceSynchronize();  // like cudaFree: nothing in flight may still touch the block when its next owner gets it
  std::lock_guard<std::mutex> lk(g_mem_mu);
  auto it = g_mem_size.find(p);
  if (!mem_cache_on() || it == g_mem_size.end()) { if (it != g_mem_size.end()) g_mem_size.erase(it); cudaFree(p); return; }
  g_mem_free[it->second.first].insert({it->second.second, p});
}

// the engines' small page-locked blocks (flag words, counters; 512 bytes each) are recycled too: cudaHostAlloc was seen to
// take up to 0.15 s right after a large page-locked region had been released
struct CachedComm { int device, rank, nranks; ncclComm_t comm; };
static std::vector<CachedComm> g_comm_cache;
static std::vector<void *> g_host_small;
static void *host_small_alloc()
{
  { std::lock_guard<std::mutex> lk(g_mem_mu); if (!g_host_small.empty()) { void *p = g_host_small.back(); g_host_small.pop_back(); memset(p, 0, 512); return p; } }
  void *p = nullptr;
  if (cudaHostAlloc(&p, 512, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  memset(p, 0, 512);
  return p;
}
static void host_small_free(void *p) { std::lock_guard<std::mutex> lk(g_mem_mu); g_host_small.push_back(p); }

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void release() { if (p) dev_free(p); p = nullptr; n = 0; }
  // grow to at least m elements; keep = number of leading elements to preserve
  void ensure(dem_engine *E, size_t m, size_t keep = 0, cudaStream_t st = 0);
};

struct ListSet {  // one ELLPACK neighbour list + history (two sets ping-pong across rebuilds)
  DevBuf<unsigned> nbr;
  DevBuf<int> ptag, numneigh;
  DevBuf<double4> hist;
  int cap = 0, maxk = 0, dnum = 0, hslots = 0, valid = 0;  // dnum = 32-byte history records per contact
  int fmt = 0;        // 0: full list, both owners hold the history (bond decks, k_step_bond); 1: owner list (dem_pairs.cuh)
  int nlocal = 0;     // owned particles when the list was built (decides who owned a pair of the OLD list)
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
  double bp[T_COUNT][MAXT + 1][MAXT + 1] = {{{0}}};  // bond tables, indexed by the T_B_* table id
  double tsCreateBond = 0.0, rmin = 0.0;
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
  // brick decomposition (comm_brick.cpp / procmap.cpp): rank -> (ix,iy,iz), sub-box, 6 neighbours
  int pgrid[3] = {1, 1, 1}, myloc[3] = {0, 0, 0}, user_grid = 0;
  double sublo[3] = {0, 0, 0}, subhi[3] = {1, 1, 1};
  ncclComm_t comm = nullptr;
  int comm_bad = 0;  // an NCCL call failed or a halo wait timed out: the communicator is not handed to a later engine
  // ghost swaps, LAMMPS order: for dim 0..2: (send to lo neighbour, send to hi neighbour)
  struct Swap {
    int dim = 0, side = 0, peer = -1, self = 0, nsend = 0, nrecv = 0, gfirst = 0; double shift = 0.0; DevBuf<int> list;
    // peer-memory path (valid between rebuilds): the receiver's record arrays, ghost offset and buffer parity
    double4 *pbase[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; int *psig = nullptr; int pgfirst = 0, pcur0 = 0, p2p = 0, serial = 0;
  };
  Swap swaps[6]; int nswap = 0;
  DevBuf<double4> sbuf, rbuf_unused;
  DevBuf<int> sbuf_i, gorder, gone, cnt_dev;
  DevBuf<double> migs, migr;
  int *hcnt = nullptr;  // pinned host scratch for counts
  DevBuf<int> order, order_keys; int order_valid = 0;  // ascending-tag permutation for read-back
  DevBuf<char> stage;  // device staging for upload / read-back
  DevBuf<int> dflag;    // device-side rebuild trigger (multi-rank: all-reduced)
  DevBuf<int> flo, fhi, slo, shi;
  // cells / sort
  GridP grid = {};
  long ncells = 0;
  DevBuf<int> ocs, oce, gcs, gce, perm, vals;
  DevBuf<unsigned> keys, keys2;
  DevBuf<char> cubtmp;
  ListSet ls[2];
  int lcur = 0;
  DevBuf<double4> res;  // owner list: per-contact result records
  DevBuf<double4> cout; long cout_serial = 0;  // option contact_output: per-contact force / torque of the last materialising evaluation
  long serial = 0;      // step launches so far (stamps the result records)
  DevBuf<int> overflow;
  DevBuf<unsigned long long> counters;
  int *hflag = nullptr;  // mapped pinned flags: [0] rebuild trigger, [1] history overflow, [2] moving-mesh trigger
  // triangle-mesh walls (dem_mesh.h)
  std::vector<MeshHost> meshes;
  struct MeshWall { std::string id; ModelP m; std::vector<int> mesh; };
  std::vector<MeshWall> mwalls;
  std::vector<TriRec> htri; std::vector<int> hcn;
  DevBuf<double> dmforce, dmpref; int any_stress = 0;  // fix mesh/surface/stress accumulators / reference points
  DevBuf<TriRec> dtri; DevBuf<int> dcn, dcell_start, dcell_tri, mint[2]; DevBuf<double> dnodes_last; DevBuf<double4> mhist[2];
  int mcur = 0, mslots = 8, mcand = 16, mhrec = 0, mesh_ready = 0, grid_ready = 0, any_moving = 0;
  double mgorg[3] = {0, 0, 0}, mginv[3] = {1, 1, 1}; int mgnc[3] = {1, 1, 1};
  long next_reneighbor = -1;
  // CUDA IPC mappings of neighbour ranks' buffers (key = the 64 handle bytes)
  struct IpcMap { unsigned char key[64]; void *ptr; int rank, gen; };
  int alloc_gen = 0;  // bumped whenever the record arrays are reallocated (their IPC handles change meaning)
  std::vector<IpcMap> ipc;
  DevBuf<int> hsig;       // [0..5] incoming halo serials per swap, [8..13] block counters of my pack kernels
  int cur0 = 0, p2p_ok = 0;
  cudaEvent_t fev[2] = {nullptr, nullptr};  // "flags of slot k are on the host"
  // step flags between ranks over peer memory (k_push / k_wait): my flag box, every rank's box, serial of the last hand-over
  // fused ghost push (fused_halo_setup): image table, block order, device parameter block; fz_on: the next launch_step uses it
  double cdf_user = 0.0;  // neigh_modify contact_distance_factor, 0 = not given
  int need_setup = 0;  // particles were inserted: the next dem_run needs a dem_setup first (lists, forces)
  std::vector<std::string> xf_id; std::vector<XForce> xf;  // fix addforce / fix viscous (dem_set_extra_force)
  int ins_mass = 0;    // the upload kernels form the mass like fix insert/* does (set by dem_insert_step_end only)
  int ins_open = 0;    // dem_insert_step_begin has run: the timestep is half done
  DevBuf<unsigned long long> bondc;  // compute bond/counter: created, broken, scratch for the total
  DevBuf<int> img_first, img_ws, img_in; DevBuf<int4> img_tab; DevBuf<ImgP> imgp; int fz_ready = 0, fz_on = 0, fz_rq[2] = {-1, -1}, slot_zeroed = 0;
  DevBuf<int> fbox; int *peer_fbox[DEM_MAXRANKS] = {nullptr}; int fbox_ready = 0, fserial = 0, fser_slot[2] = {0, 0}, fpeer[2] = {0, 0};
  int *hflag_dev = nullptr; int fev_peer[2] = {0, 0};  // slot's flags arrive through k_wait (serial in hflag[16 + slot]) instead of memcpy + event
  int fslot = 0;                            // slot the next step writes
  const int *gate = nullptr; int gate_mask = 0;  // gate of the step being launched (nullptr: not speculative)
  // state
  int uploaded = 0, setup_done = 0, forces_valid = 0;
  int dirty = 0;  // a deck setting (box, neighbor, property, timestep-independent tables) changed after the first setup: re-derive
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
  CK(dev_alloc((void **)&q, m * sizeof(T)));
  if (keep && p) CK(cudaMemcpyAsync(q, p, std::min(keep, n) * sizeof(T), cudaMemcpyDeviceToDevice, st));
  if (p) { CK(cudaStreamSynchronize(st)); dev_free(p); }
  p = q; n = m;
}

#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) { E->comm_bad = 1; dem_fail(E, DEM_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r__)); } \
  } while (0)
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
    int cc_major = 0, cc_minor = 0;  // (cudaGetDeviceProperties takes tens of milliseconds; two attributes do not)
    CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    CK(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
    if (cc_major < 10) dem_fail(e, DEM_ERR_CUDA, "device %d is sm_%d%d; this library only contains sm_100a code", device, cc_major, cc_minor);
    CK(cudaSetDevice(device));
    e->device = device; e->rank = rank; e->nranks = nranks; e->stream = (cudaStream_t)stream;
    if (nranks < 1 || rank < 0 || rank >= nranks) dem_fail(e, DEM_ERR_ARG, "bad rank/nranks %d/%d", rank, nranks);
    if (nranks > 1) {
      if (!g_nccl.load()) dem_fail(e, DEM_ERR_CUDA, "could not load libnccl.so.2");
      // Communicator reuse is explicit: nccl_id == NULL asks for the communicator an earlier engine of this process with the
      // same (device, rank, nranks) released in good standing (ncclCommInitRank over 8 GPUs costs seconds; an SPMD job that
      // creates its engines in the same order on every rank may reuse it).  A non-NULL id always builds a fresh communicator
      // for exactly the peer group that shares the id, and retires a cached one.
      {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        for (auto it = g_comm_cache.begin(); it != g_comm_cache.end(); ++it)
          if (it->device == device && it->rank == rank && it->nranks == nranks) { e->comm = it->comm; g_comm_cache.erase(it); break; }
      }
      if (!nccl_id) {
        if (!e->comm) dem_fail(e, DEM_ERR_ARG, "nranks > 1 needs the shared 128-byte ncclUniqueId (dem_nccl_unique_id on rank 0); NULL only reuses a communicator released by an earlier engine of this process");
      } else {
        if (e->comm) { g_nccl.CommDestroy(e->comm); e->comm = nullptr; }
        ncclUniqueId id; memcpy(&id, nccl_id, sizeof id);
        NK(g_nccl.CommInitRank(&e->comm, nranks, id, rank));
      }
    }
    e->hcnt = (int *)host_small_alloc(); e->hflag = (int *)host_small_alloc();
    if (!e->hcnt || !e->hflag) dem_fail(e, DEM_ERR_CUDA, "cudaHostAlloc failed");
    for (int k = 0; k < 8; k++) e->hflag[k] = 0;
    e->pm.tdamp = 1;
  } catch (const DemFail &f) { return f.code; }
  return DEM_OK;
}

extern "C" int dem_nccl_unique_id(void *out128)
{
  if (!out128 || !g_nccl.load()) return DEM_ERR_CUDA;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return DEM_ERR_CUDA;
  memcpy(out128, &id, sizeof id);
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
  e->order.release(); e->order_keys.release(); e->stage.release(); e->valid_tmp.release(); e->wlist.release(); e->fw.release(); e->sbuf.release(); e->sbuf_i.release(); e->gorder.release(); e->gone.release(); e->cnt_dev.release(); e->migs.release(); e->migr.release(); e->dflag.release(); for (auto &sw : e->swaps) sw.list.release();
  e->flo.release(); e->fhi.release(); e->slo.release(); e->shi.release(); e->ocs.release(); e->oce.release();
  e->gcs.release(); e->gce.release(); e->perm.release(); e->vals.release(); e->keys.release(); e->keys2.release();
  e->cubtmp.release(); e->overflow.release(); e->counters.release(); e->res.release(); e->cout.release();
  e->dmforce.release(); e->dmpref.release();
  e->dtri.release(); e->dcn.release(); e->dcell_start.release(); e->dcell_tri.release(); e->dnodes_last.release();
  for (int s = 0; s < 2; s++) { e->mint[s].release(); e->mhist[s].release(); }
  for (int s = 0; s < 2; s++) { e->ls[s].nbr.release(); e->ls[s].ptag.release(); e->ls[s].numneigh.release(); e->ls[s].hist.release(); }
  for (auto &ev : e->ev) cudaEventDestroy(ev);
  if (e->hflag) host_small_free(e->hflag);
  for (int k = 0; k < 2; k++) if (e->fev[k]) cudaEventDestroy(e->fev[k]);
  e->hsig.release(); e->fbox.release(); e->bondc.release(); e->img_first.release(); e->img_tab.release(); e->imgp.release(); e->img_ws.release(); e->img_in.release();
  if (e->hcnt) host_small_free(e->hcnt);
  if (e->comm) {
    if (e->comm_bad || getenv("DEM_B200_NO_COMM_CACHE")) g_nccl.CommDestroy(e->comm);
    else { std::lock_guard<std::mutex> lk(g_mem_mu); g_comm_cache.push_back({e->device, e->rank, e->nranks, e->comm}); }
  }
  delete e;
}

extern "C" long dem_trim_memory(void) { return (long)mem_trim(-1); }

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
  e->dirty = 1;
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
  e->pgrid[0] = px; e->pgrid[1] = py; e->pgrid[2] = pz; e->user_grid = 1;
  API_END
}
extern "C" int dem_set_neighbor(dem_engine *e, double skin, int every, int delay, int check)
{
  API_BEGIN
  if (skin < 0 || every < 1 || delay < 0) dem_fail(e, DEM_ERR_ARG, "bad neighbor settings");
  if (skin != e->skin) e->dirty = 1;  // cell grid, cutneighmax, border slabs and the mesh grid depend on the skin
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

// property/global names of the two bond models -> table id
static int bond_table_of(const std::string &nm)
{
  static const std::map<std::string, int> m = {
    {"radiusMultiplierBond", T_B_LAMBDA}, {"normalBondStiffnessPerUnitArea", T_B_KN}, {"tangentialBondStiffnessPerUnitArea", T_B_KT},
    {"dampingNormalForceBond", T_B_DFN}, {"dampingTangentialForceBond", T_B_DFT}, {"dampingNormalTorqueBond", T_B_DTN}, {"dampingTangentialTorqueBond", T_B_DTT},
    {"maxDistanceBond", T_B_MAXDIST}, {"maxSigmaBond", T_B_MAXSIGMA}, {"maxTauBond", T_B_MAXTAU}, {"createDistanceBond", T_B_CREATEDIST}, {"ratioTensionCompression", T_B_RATIOTC},
    {"radiusMultiplierBondnonlinear", T_B_LAMBDA}, {"dampingNormalForceBondnonlinear", T_B_DFN}, {"dampingTangentialForceBondnonlinear", T_B_DFT},
    {"dampingNormalTorqueBondnonlinear", T_B_DTN}, {"dampingTangentialTorqueBondnonlinear", T_B_DTT}, {"maxDistanceBondnonlinear", T_B_MAXDIST},
    {"maxSigmaBondnonlinear", T_B_MAXSIGMA}, {"maxTauBondnonlinear", T_B_MAXTAU}, {"createDistanceBondnonlinear", T_B_CREATEDIST},
    {"ratioTensionCompressionBondnonlinear", T_B_RATIOTC}, {"stiffnessPerUnitAreaK_fn1", T_B_K_FN1}, {"stiffnessPerUnitAreaKu_fn1", T_B_KU_FN1},
    {"stiffnessPerUnitAreaKc_fn1", T_B_KC_FN1}, {"stiffnessPerUnitAreaK_fn2", T_B_K_FN2}, {"stiffnessPerUnitAreaKu_fn2", T_B_KU_FN2}, {"stiffnessPerUnitAreaKc_fn2", T_B_KC_FN2},
    {"stiffnessPerUnitAreaK_ft", T_B_K_FT}, {"stiffnessPerUnitAreaK_tn", T_B_K_TN}, {"stiffnessPerUnitAreaKu_tn", T_B_KU_TN}, {"stiffnessPerUnitAreaKc_tn", T_B_KC_TN},
    {"stiffnessPerUnitAreaK_tt", T_B_K_TT}, {"stiffnessPerUnitAreaKu_tt", T_B_KU_TT}, {"stiffnessPerUnitAreaKc_tt", T_B_KC_TT},
    // normal models hysteretic/nonlinear1|2
    {"LoadingStiffness", T_H_KEL}, {"UnloadingStiffness", T_H_KN2K1}, {"coefficientAdhesionStiffness", T_H_KN2KC}, {"coefficientPlasticityDepth", T_H_PHIF},
    {"pullOffForce", T_H_FADH}, {"alphaCustom", T_H_ALPHA}, {"cinCustom", T_H_CIN}, {"aoneCustom", T_H_A1}, {"atwoCustom", T_H_A2}, {"athreeCustom", T_H_A3},
    {"kcinCustom", T_H_KCIN},
    {"dissipationNormalForceBond", T_D_FN}, {"dissipationTangentialForceBond", T_D_FT}, {"dissipationNormalTorqueBond", T_D_TN}, {"dissipationTangentialTorqueBond", T_D_TT}};
  auto it = m.find(nm);
  return it == m.end() ? -1 : it->second;
}

extern "C" int dem_set_contact_distance_factor(dem_engine *e, double f)
{
  API_BEGIN
  if (!(f >= 1.0)) dem_fail(e, DEM_ERR_ARG, "Illegal neigh_modify command. Please set contact_distance_factor value >=1");
  if (f >= 10.0) dem_fail(e, DEM_ERR_ARG, "contactDistanceFactor must be < 10");
  e->cdf_user = f; e->dirty = 1; e->ls[0].valid = e->ls[1].valid = 0;
  API_END
}

extern "C" int dem_set_property(dem_engine *e, const char *name, const char *kind, const double *v, int n)
{
  API_BEGIN
  const int T = e->ntypes;
  std::string nm(name), kd(kind);
  if (kd == "scalar") {
    if (nm == "characteristicVelocity" && n == 1) e->charVel = v[0];
    else if ((nm == "tsCreateBond" || nm == "tsCreateBondnonlinear") && n == 1) e->tsCreateBond = v[0];
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
    if (!dst) { const int b = bond_table_of(nm); if (b >= 0) dst = e->bp[b]; }
    if (!dst) dem_fail(e, DEM_ERR_UNSUPPORTED, "peratomtypepair property %s not on the hot path", name);
    for (int i = 0; i < T; i++) for (int j = 0; j < T; j++) {
      if (v[i * T + j] != v[j * T + i]) dem_fail(e, DEM_ERR_ARG, "%s: per-atomtype property matrix must be symmetric", name);
      dst[i + 1][j + 1] = v[i * T + j];
    }
  } else dem_fail(e, DEM_ERR_ARG, "unknown property kind %s", kind);
  e->have_prop[nm] = 1;
  e->dirty = 1;  // the material tables are re-derived by the next dem_setup
  API_END
}

// model selection in the reference's fixed keyword order (contact_models.cpp:158-260)
static void parse_model_select(dem_engine *e, int &argc, const char *const *&a, ModelP &m)
{
  memset(&m, 0, sizeof m); m.tdamp = 1; m.off_shear = m.off_roll = -1;
  if (argc > 1 && !strcmp(a[0], "model")) {
    if (!strcmp(a[1], "hertz")) m.normal = N_HERTZ;
    else if (!strcmp(a[1], "hooke")) m.normal = N_HOOKE;
    else if (!strcmp(a[1], "hysteretic/nonlinear1")) { m.normal = N_HYST1; m.limitForce = 1; }  // limitForce defaults to on (normal_model_hysteretic_nonlinear1.h:100)
    else if (!strcmp(a[1], "hysteretic/nonlinear2")) { m.normal = N_HYST2; m.limitForce = 1; }
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "normal model '%s' is outside the hot-path scope (hertz, hooke, hysteretic/nonlinear1, hysteretic/nonlinear2)", a[1]);
    a += 2; argc -= 2;
  } else dem_fail(e, DEM_ERR_ARG, "expected 'model <normal model>'");
  if (argc > 1 && !strcmp(a[0], "tangential")) {
    if (!strcmp(a[1], "history")) m.tangential = 1;
    else if (!strcmp(a[1], "off")) m.tangential = 0;
    else if (!strcmp(a[1], "hysteretic/nonlinear")) {  // reads sidata.deltaZero, which only the hysteretic normal laws set (tangential_model_hysteretic_nonlinear.h:150)
      if (m.normal < N_HYST1) dem_fail(e, DEM_ERR_UNSUPPORTED, "tangential model hysteretic/nonlinear needs normal model hysteretic/nonlinear1|2");
      m.tangential = 2;
    }
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "tangential model '%s' is outside the hot-path scope (history, hysteretic/nonlinear)", a[1]);
    a += 2; argc -= 2;
  }
  m.tension = m.compression = m.shearf = m.ntorque = m.ttorque = m.damping = 1;
  if (argc > 1 && !strcmp(a[0], "cohesion")) {
    if (!strcmp(a[1], "bond")) m.cohesion = C_BOND; else if (!strcmp(a[1], "bond/nonlinear")) m.cohesion = C_BONDNL;
    else if (!strcmp(a[1], "off")) m.cohesion = C_OFF;
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "cohesion model '%s' is outside the hot-path scope (bond, bond/nonlinear)", a[1]);
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "rolling_friction")) {
    if (!strcmp(a[1], "cdt")) m.rolling = R_CDT; else if (!strcmp(a[1], "cdtnonlinear2")) { m.rolling = R_CDT; m.cdtnl2 = 1; } else if (!strcmp(a[1], "epsd")) m.rolling = R_EPSD;
    else if (!strcmp(a[1], "epsd2")) m.rolling = R_EPSD2; else if (!strcmp(a[1], "off")) m.rolling = R_OFF;
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "rolling model '%s' is outside the hot-path scope (cdt, cdtnonlinear2, epsd, epsd2)", a[1]);
    a += 2; argc -= 2;
  }
  if (argc > 1 && !strcmp(a[0], "surface")) {
    if (strcmp(a[1], "default")) dem_fail(e, DEM_ERR_UNSUPPORTED, "surface model '%s' is outside the hot-path scope", a[1]);
    a += 2; argc -= 2;
  }
  if ((m.rolling == R_EPSD || m.rolling == R_EPSD2) && !m.tangential)
    dem_fail(e, DEM_ERR_ARG, "rolling_friction epsd/epsd2 requires tangential history");
  m.dnum = 0; m.hrec = 0; m.rec_shear = m.rec_roll = m.rec_bond = m.rec_norm = -1; m.off_bond = m.off_norm = -1;
  if (m.normal == N_HYST1 || m.normal == N_HYST2) {  // the normal model's 12 history values come first (model construction order)
    if (m.cohesion) dem_fail(e, DEM_ERR_UNSUPPORTED, "normal model hysteretic/nonlinear with a bond model is outside the hot-path scope");
    m.off_norm = m.dnum; m.dnum += 12; m.rec_norm = m.hrec; m.hrec += 3;
  }
  if (m.cohesion) {  // history slot order = model construction order: cohesion, tangential, rolling (contact_models.h:141-145)
    m.nbond = m.cohesion == C_BOND ? 14 : 28; m.off_bond = 0; m.dnum += m.nbond;
    m.rec_bond = 0; m.nbrec = (m.nbond + 1 + 3) / 4; m.hrec += m.nbrec;
  }
  if (m.tangential) { m.off_shear = m.dnum; m.dnum += m.tangential == 2 ? 7 : 3; m.rec_shear = m.hrec++; }  // (hysteretic/nonlinear: shear xyz | shrmag_0 in one record; values 4..6 stay 0)
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
    else if (m.cohesion && !strcmp(a[0], "stressBreak")) m.stressBreak = on;
    else if (m.cohesion && !strcmp(a[0], "tensionStress")) m.tension = on;
    else if (m.cohesion && !strcmp(a[0], "compressionStress")) m.compression = on;
    else if (m.cohesion && !strcmp(a[0], "shearStress")) m.shearf = on;
    else if (m.cohesion && !strcmp(a[0], "normalTorqueStress")) m.ntorque = on;
    else if (m.cohesion && !strcmp(a[0], "shearTorqueStress")) m.ttorque = on;
    else if (m.cohesion && !strcmp(a[0], "createBondAlways")) m.createAlways = on;
    else if (m.cohesion && !strcmp(a[0], "dampingBond")) m.damping = on;
    else if (m.cohesion == C_BOND && !strcmp(a[0], "dissipationBond")) m.dissipation = on;
    else if (m.cohesion && !strcmp(a[0], "dampingBondSmooth")) { m.dampingSmooth = on; if (on) m.damping = 1; }
    else if (m.cohesion == C_BOND && !strcmp(a[0], "ratioTensionCompression")) m.ratioTC = on;
    else if (m.cohesion == C_BONDNL && !strcmp(a[0], "ratioTensionCompressionBond")) m.ratioTC = on;
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
  if (W.p.m.normal >= N_HYST1) dem_fail(e, DEM_ERR_UNSUPPORTED, "normal model hysteretic/nonlinear on a wall is outside the hot-path scope (pair style only)");
  if (W.p.m.cohesion) dem_fail(e, DEM_ERR_UNSUPPORTED, "bond models on walls are outside the hot-path scope");
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


// ------------------------------------------------------------------------------------------------
// triangle-mesh walls
extern "C" int dem_add_mesh(dem_engine *e, const char *id, int atom_type, const double *nodes9, long ntri, int argc, const char *const *argv)
{
  API_BEGIN
  if (e->setup_done) dem_fail(e, DEM_ERR_STATE, "meshes cannot be added after setup");
  if ((int)e->meshes.size() == DEM_MAXMESH) dem_fail(e, DEM_ERR_OVERFLOW, "at most %d meshes", DEM_MAXMESH);
  if (!id || !nodes9 || ntri < 1) dem_fail(e, DEM_ERR_ARG, "mesh needs an id and at least one triangle");
  if (atom_type < 1 || atom_type > e->ntypes) dem_fail(e, DEM_ERR_ARG, "1 <= type <= max type as defined in create_box");
  for (auto &m : e->meshes) if (m.id == id) dem_fail(e, DEM_ERR_ARG, "fix id %s already in use", id);
  MeshHost M;
  M.id = id; M.atom_type = atom_type; M.ntri = (int)ntri;
  M.nodes.assign(nodes9, nodes9 + 9 * ntri);
  for (int k = 0; k < argc; k += 2) {
    if (k + 1 >= argc) dem_fail(e, DEM_ERR_ARG, "mesh keyword '%s' needs a value", argv[k]);
    if (!strcmp(argv[k], "curvature")) M.curvature = cos(atof(argv[k + 1]) * 3.14159265358979323846 / 180.);
    else if (!strcmp(argv[k], "precision")) M.precision = atof(argv[k + 1]);
    else if (!strcmp(argv[k], "stress")) M.stress = !strcmp(argv[k + 1], "on");  // fix mesh/surface/stress (mesh_module_stress.cpp:120-140)
    else if (!strcmp(argv[k], "reference_point")) {
      if (k + 3 >= argc) dem_fail(e, DEM_ERR_ARG, "mesh keyword 'reference_point' needs three values");
      for (int d = 0; d < 3; d++) M.p_ref[d] = atof(argv[k + 1 + d]);
      k += 2;
    }
    else dem_fail(e, DEM_ERR_UNSUPPORTED, "fix mesh/surface keyword '%s' is outside the hot-path scope (apply scale/move/rotate to the nodes before the call)", argv[k]);
  }
  e->meshes.push_back(M);
  API_END
}
extern "C" int dem_move_mesh(dem_engine *e, const char *mesh_id, int argc, const char *const *argv)
{
  API_BEGIN
  // may also arrive between two runs (the t01a tutorial deck starts its mesh after the settling run): the next dem_setup
  // rebuilds the lists with the moving-mesh candidate skin and the mesh starts to move with the first step
  for (auto &m : e->meshes) if (m.id == mesh_id) {
    if (m.moving) dem_fail(e, DEM_ERR_UNSUPPORTED, "one fix move/mesh per mesh (superposed movers are outside the hot-path scope)");
    if (argc >= 1 && !strcmp(argv[0], "linear")) {
      if (argc != 4) dem_fail(e, DEM_ERR_ARG, "Not enough arguments for movement type linear");
      for (int d = 0; d < 3; d++) m.vel[d] = atof(argv[1 + d]);
      m.moving = 1;
    } else if (argc >= 1 && !strcmp(argv[0], "rotate")) {  // mesh_mover_rotation.cpp:58-82
      if (argc < 11) dem_fail(e, DEM_ERR_ARG, "Not enough arguments for movement type rotate");
      if (strcmp(argv[1], "origin")) dem_fail(e, DEM_ERR_ARG, "Expected keyword 'origin'");
      if (strcmp(argv[5], "axis")) dem_fail(e, DEM_ERR_ARG, "Expected keyword 'axis'");
      if (strcmp(argv[9], "period")) dem_fail(e, DEM_ERR_ARG, "Expected keyword 'period'");
      for (int d = 0; d < 3; d++) { m.rot_origin[d] = atof(argv[2 + d]); m.rot_axis[d] = atof(argv[6 + d]); }
      const double norm = sqrt(m.rot_axis[0] * m.rot_axis[0] + m.rot_axis[1] * m.rot_axis[1] + m.rot_axis[2] * m.rot_axis[2]);
      const double invnorm = (norm == 0.) ? 0. : 1. / norm;  // vectorNormalize3D, vector_liggghts.h:63-70
      if (norm == 0.) dem_fail(e, DEM_ERR_ARG, "fix move/mesh rotate: axis = 0");
      for (int d = 0; d < 3; d++) m.rot_axis[d] *= invnorm;
      m.rot_omega = 2. * 3.14159265358979323846 / atof(argv[10]);
      m.moving = 2;
    } else dem_fail(e, DEM_ERR_UNSUPPORTED, "fix move/mesh style '%s' is outside the hot-path scope (linear, rotate)", argc ? argv[0] : "");
    e->any_moving = 1;
    return DEM_OK;
  }
  dem_fail(e, DEM_ERR_ARG, "no mesh with id %s", mesh_id);
  API_END
}
extern "C" int dem_add_wall_mesh(dem_engine *e, const char *id, int argc, const char *const *argv)
{
  API_BEGIN
  if (e->setup_done) dem_fail(e, DEM_ERR_STATE, "walls cannot be added after setup");
  dem_engine::MeshWall W; W.id = id;
  parse_model_select(e, argc, argv, W.m);
  if (W.m.normal >= N_HYST1) dem_fail(e, DEM_ERR_UNSUPPORTED, "normal model hysteretic/nonlinear on a mesh wall is outside the hot-path scope (pair style only)");
  if (W.m.cohesion) dem_fail(e, DEM_ERR_UNSUPPORTED, "bond models on walls are outside the hot-path scope");
  if (argc < 4 || strcmp(argv[0], "mesh")) dem_fail(e, DEM_ERR_ARG, "Need to use define style 'mesh' or 'primitive'");
  if (strcmp(argv[1], "n_meshes")) dem_fail(e, DEM_ERR_ARG, "have to define 'n_meshes' before 'meshes'");
  const int nm = atoi(argv[2]);
  if (nm < 1) dem_fail(e, DEM_ERR_ARG, "'n_meshes' > 0 required");
  if (argc < 4 + nm || strcmp(argv[3], "meshes")) dem_fail(e, DEM_ERR_ARG, "Need to provide the number and a list of meshes by using 'n_meshes' and 'meshes'");
  for (int k = 0; k < nm; k++) {
    int found = -1;
    for (size_t m = 0; m < e->meshes.size(); m++) if (e->meshes[m].id == argv[4 + k]) found = (int)m;
    if (found < 0) dem_fail(e, DEM_ERR_ARG, "could not find fix mesh id %s", argv[4 + k]);
    if (e->meshes[found].wall >= 0) dem_fail(e, DEM_ERR_ARG, "mesh %s is already used by another wall", argv[4 + k]);
    e->meshes[found].wall = (int)e->mwalls.size();
    W.mesh.push_back(found);
  }
  argv += 4 + nm; argc -= 4 + nm;
  parse_model_settings(e, argc, argv, W.m);
  e->mwalls.push_back(W);
  API_END
}

static bool have_mesh_walls(const dem_engine *E) { return !E->mwalls.empty(); }

static MeshP mesh_params(dem_engine *E)
{
  MeshP M;
  memset(&M, 0, sizeof M);
  M.ntri = (int)E->htri.size(); M.nmesh = (int)E->meshes.size(); M.mslots = E->mslots; M.mcand = E->mcand; M.cap = E->cap; M.hrec = E->mhrec;
  M.mforce = E->any_stress ? E->dmforce.p : nullptr; M.mpref = E->any_stress ? E->dmpref.p : nullptr;
  M.tri = E->dtri.p; M.nodes_last = E->dnodes_last.p; M.cn = E->dcn.p; M.cell_start = E->dcell_start.p; M.cell_tri = E->dcell_tri.p;
  for (int d = 0; d < 3; d++) { M.gorg[d] = E->mgorg[d]; M.ginv[d] = E->mginv[d]; M.gnc[d] = E->mgnc[d]; }
  M.mint = E->mint[E->mcur].p; M.mhist = E->mhist[E->mcur].p;
  for (size_t m = 0; m < E->meshes.size(); m++) {
    const MeshHost &H = E->meshes[m];
    MeshMeta &mm = M.meta[m];
    mm.stress = H.stress;
    mm.atom_type = H.atom_type; mm.wall = H.wall; mm.moving = H.moving; mm.first = H.first; mm.ntri = H.ntri; mm.precision = H.precision;
    for (int d = 0; d < 3; d++) mm.vel[d] = H.vel[d];
    if (H.moving == 2) {  // MultiNodeMesh::rotate(dAngle, axis, p), multi_node_mesh_I.h:620-640
      const double dphi = H.rot_omega * E->dt;
      double ax[3] = {H.rot_axis[0], H.rot_axis[1], H.rot_axis[2]};
      const double sinv = 1. / sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      for (int d = 0; d < 3; d++) ax[d] = sinv * ax[d];
      mm.rot_dq[0] = cos(dphi * 0.5);
      for (int d = 0; d < 3; d++) { mm.rot_dq[d + 1] = ax[d] * sin(dphi * 0.5); mm.rot_origin[d] = H.rot_origin[d]; mm.rot_omegavec[d] = H.rot_axis[d] * H.rot_omega; }
      mm.rot_trans = (H.rot_origin[0] * H.rot_origin[0] + H.rot_origin[1] * H.rot_origin[1] + H.rot_origin[2] * H.rot_origin[2]) > 0.;
    }
    if (H.wall >= 0) M.wm[m] = E->mwalls[H.wall].m;
  }
  M.overflow = E->overflow.p;
  return M;
}

// setup-time: geometry + topology on the host, triangle records to the device, per-particle rows
static void mesh_prepare(dem_engine *E)
{
  if (E->mesh_ready || !have_mesh_walls(E)) return;
  E->htri.clear(); E->hcn.clear(); E->mhrec = 0; E->any_moving = 0;
  for (size_t m = 0; m < E->meshes.size(); m++) {
    MeshHost &H = E->meshes[m];
    H.first = (int)E->htri.size();
    const std::string err = meshhost::derive(E->lo, E->hi, H, (int)m, E->htri, E->hcn);
    if (!err.empty()) dem_fail(E, DEM_ERR_ARG, "%s", err.c_str());
    if (H.wall >= 0) E->mhrec = std::max(E->mhrec, E->mwalls[H.wall].m.hrec);
    if (H.moving) E->any_moving = 1;
  }
  {  // coplanar node-neighbour lists: (count, entries...) per triangle -> CSR [ntri+1 offsets into the same array][entries]
    const size_t T = E->htri.size();
    std::vector<int> csr(T + 1, 0), ent;
    size_t r = 0;
    for (size_t t = 0; t < T; t++) { const int c = E->hcn[r++]; csr[t] = (int)(T + 1 + ent.size()); ent.insert(ent.end(), E->hcn.begin() + r, E->hcn.begin() + r + c); r += c; }
    csr[T] = (int)(T + 1 + ent.size());
    csr.insert(csr.end(), ent.begin(), ent.end());
    E->hcn.swap(csr);
  }
  if (E->opt.count("meshslots")) E->mslots = std::min(30, std::max(2, (int)E->opt["meshslots"]));
  if (E->opt.count("meshcand")) E->mcand = std::max(4, (int)E->opt["meshcand"]);
  cudaStream_t st = E->stream;
  E->any_stress = 0;
  for (auto &H : E->meshes) if (H.stress) E->any_stress = 1;
  if (E->any_stress) {
    E->dmforce.ensure(E, 6 * DEM_MAXMESH); E->dmpref.ensure(E, 3 * DEM_MAXMESH);
    double hp[3 * DEM_MAXMESH] = {0};
    for (size_t m = 0; m < E->meshes.size(); m++) for (int d = 0; d < 3; d++) hp[3 * m + d] = E->meshes[m].p_ref[d];
    CK(cudaMemcpyAsync(E->dmpref.p, hp, sizeof hp, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(E->dmforce.p, 0, 6 * DEM_MAXMESH * sizeof(double), st));
    CK(cudaStreamSynchronize(st));
  }
  E->dtri.ensure(E, E->htri.size()); E->dcn.ensure(E, E->hcn.size()); E->dnodes_last.ensure(E, 9 * E->htri.size());
  CK(cudaMemcpyAsync(E->dtri.p, E->htri.data(), E->htri.size() * sizeof(TriRec), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(E->dcn.p, E->hcn.data(), E->hcn.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  const size_t rows = 1 + E->mslots + E->mcand;
  for (int b = 0; b < 2; b++) {
    E->mint[b].release(); E->mint[b].ensure(E, rows * E->cap);
    CK(cudaMemsetAsync(E->mint[b].p, 0xFF, rows * E->cap * sizeof(int), st));   // partner rows: -1 = free
    CK(cudaMemsetAsync(E->mint[b].p, 0, (size_t)E->cap * sizeof(int), st));     // row 0: candidate count
    E->mhist[b].release();
    if (E->mhrec) { E->mhist[b].ensure(E, (size_t)E->mslots * E->mhrec * E->cap); CK(cudaMemsetAsync(E->mhist[b].p, 0, (size_t)E->mslots * E->mhrec * E->cap * sizeof(double4), st)); }
  }
  CK(cudaStreamSynchronize(st));
  E->mcur = 0; E->mesh_ready = 1; E->grid_ready = 0; E->next_reneighbor = -1;
}

// coarse triangle grid (moving meshes: from the current device nodes at every rebuild)
static void mesh_grid(dem_engine *E)
{
  if (E->grid_ready && !E->any_moving) return;
  cudaStream_t st = E->stream;
  if ((E->grid_ready || E->setup_done) && E->any_moving) {  // (setup_done: the grid was invalidated by a changed skin)
    CK(cudaMemcpyAsync(E->htri.data(), E->dtri.p, E->htri.size() * sizeof(TriRec), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // moving meshes: edge vectors / normals / obtuse index are recomputed from the nodes at every rebuild, as the
    // reference does (SurfaceMesh::refreshOwned); the records travel to the host for the grid anyway
    for (size_t m = 0; m < E->meshes.size(); m++) if (E->meshes[m].moving) {
      const MeshHost &H = E->meshes[m];
      for (int t = 0; t < H.ntri; t++) meshhost::surf_refresh(E->htri[H.first + t]);
      CK(cudaMemcpyAsync(E->dtri.p + H.first, E->htri.data() + H.first, (size_t)H.ntri * sizeof(TriRec), cudaMemcpyHostToDevice, st));
    }
  }
  std::vector<int> cs, ct;
  meshhost::build_grid(E->lo, E->hi, 4.0 * E->cutneighmax, E->cutneighmax + E->skin, E->htri, E->mgorg, E->mginv, E->mgnc, cs, ct);
  E->dcell_start.ensure(E, cs.size()); E->dcell_tri.ensure(E, std::max<size_t>(ct.size(), 1));
  CK(cudaMemcpyAsync(E->dcell_start.p, cs.data(), cs.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!ct.empty()) CK(cudaMemcpyAsync(E->dcell_tri.p, ct.data(), ct.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  E->grid_ready = 1;
}

// rebuild: carry the per-particle rows through the sort permutation, then candidates + contact-row maintenance
static void mesh_rebuild(dem_engine *E, int n, bool permuted)
{
  cudaStream_t st = E->stream;
  mesh_grid(E);
  const int src = E->mcur, dst = E->mcur ^ 1;
  for (int attempt = 0; attempt < 6; attempt++) {
    const size_t rows = 1 + E->mslots + E->mcand;
    E->mint[dst].ensure(E, rows * E->cap);
    if (n) {
      if (permuted) {
        k_gather_rows<int><<<GRID(n, 256), 256, 0, st>>>(n, 1 + E->mslots, (size_t)E->cap, (size_t)E->cap, E->perm.p, E->mint[src].p, E->mint[dst].p);
        if (E->mhrec) k_gather_rows<double4><<<GRID(n, 256), 256, 0, st>>>(n, E->mslots * E->mhrec, (size_t)E->cap, (size_t)E->cap, E->perm.p, E->mhist[src].p, E->mhist[dst].p);
      } else {
        CK(cudaMemcpyAsync(E->mint[dst].p, E->mint[src].p, (size_t)(1 + E->mslots) * E->cap * sizeof(int), cudaMemcpyDeviceToDevice, st));
        if (E->mhrec) CK(cudaMemcpyAsync(E->mhist[dst].p, E->mhist[src].p, (size_t)E->mslots * E->mhrec * E->cap * sizeof(double4), cudaMemcpyDeviceToDevice, st));
      }
      E->launches += 2;
    }
    E->mcur = dst;
    CK(cudaMemsetAsync(E->overflow.p, 0, 2 * sizeof(int), st));
    MeshP M = mesh_params(E);
    mesh_launch_candidates(M, n, E->xr[E->cur].p, E->skin, E->cdf, st);
    E->launches++;
    int ov[2] = {0, 0};
    CK(cudaMemcpyAsync(ov, E->overflow.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ov[0] <= E->mcand) break;
    if (attempt == 5) dem_fail(E, DEM_ERR_OVERFLOW, "a particle has %d candidate triangles", ov[0]);
    E->mcand = ov[0] + 4; E->mcur = src;
  }
  // the other buffer must be able to take the rows at the next rebuild
  E->mint[src].ensure(E, (size_t)(1 + E->mslots + E->mcand) * E->cap);
  MeshP M = mesh_params(E);
  mesh_launch_hold(M, st);
  E->launches++;
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
// fix addforce (kind 0: constant fx fy fz) / fix viscous (kind 1: gamma) on a group; a fix id that exists is replaced, n < 0 removes it
extern "C" int dem_set_extra_force(dem_engine *e, const char *id, int kind, int groupbit, const double *values, int n)
{
  API_BEGIN
  if (!id || (kind != 0 && kind != 1)) dem_fail(e, DEM_ERR_ARG, "extra force: kind 0 (addforce) or 1 (viscous)");
  size_t k = 0;
  for (; k < e->xf_id.size(); k++) if (e->xf_id[k] == id) break;
  if (n < 0) { if (k == e->xf_id.size()) dem_fail(e, DEM_ERR_ARG, "Could not find fix ID %s to delete", id); e->xf_id.erase(e->xf_id.begin() + k); e->xf.erase(e->xf.begin() + k); return DEM_OK; }
  if (n != (kind == 0 ? 3 : 1) || !values) dem_fail(e, DEM_ERR_ARG, "extra force: addforce takes 3 values, viscous 1");
  XForce X; X.kind = kind; X.bit = groupbit; X.v[0] = values[0]; X.v[1] = n > 1 ? values[1] : 0.0; X.v[2] = n > 2 ? values[2] : 0.0;
  if (k == e->xf_id.size()) {
    if (e->xf.size() >= DEM_MAXXF) dem_fail(e, DEM_ERR_UNSUPPORTED, "more than %d fix addforce / viscous", DEM_MAXXF);
    e->xf_id.push_back(id); e->xf.push_back(X);
  } else e->xf[k] = X;
  e->forces_valid = 0;
  API_END
}
extern "C" int dem_set_integrate(dem_engine *e, int bit) { API_BEGIN e->integbit = bit; API_END }

// ------------------------------------------------------------------------------------------------
static void ensure_particle_cap(dem_engine *E, long need, long keep)
{
  if (need <= E->cap) return;
  const int oldcap = E->cap;
  E->alloc_gen++;
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
  if (E->mesh_ready && oldcap) {  // mesh rows: [rows][cap] int / double4
    for (int b = 0; b < 2; b++) {
      const int rows = 1 + E->mslots + E->mcand;
      DevBuf<int> ni; ni.ensure(E, (size_t)rows * ncap, 0, st);
      CK(cudaMemsetAsync(ni.p, 0xFF, (size_t)rows * ncap * sizeof(int), st)); CK(cudaMemsetAsync(ni.p, 0, (size_t)ncap * sizeof(int), st));
      if (E->mint[b].p && keep) CK(cudaMemcpy2DAsync(ni.p, ncap * sizeof(int), E->mint[b].p, (size_t)oldcap * sizeof(int), std::min<long>(keep, oldcap) * sizeof(int), std::min<size_t>(rows, E->mint[b].n / oldcap), cudaMemcpyDeviceToDevice, st));
      CK(cudaStreamSynchronize(st)); E->mint[b].release(); E->mint[b] = ni;
      if (E->mhrec) {
        const int hr = E->mslots * E->mhrec;
        DevBuf<double4> nh; nh.ensure(E, (size_t)hr * ncap, 0, st);
        CK(cudaMemsetAsync(nh.p, 0, (size_t)hr * ncap * sizeof(double4), st));
        if (E->mhist[b].p && keep) CK(cudaMemcpy2DAsync(nh.p, ncap * sizeof(double4), E->mhist[b].p, (size_t)oldcap * sizeof(double4), std::min<long>(keep, oldcap) * sizeof(double4), hr, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st)); E->mhist[b].release(); E->mhist[b] = nh;
      }
    }
  }
  if (E->nwrows) { E->whist_tmp.release(); E->whist_tmp.ensure(E, (size_t)E->nwrows * ncap, 0, st); }
  E->cap = (int)ncap;
}


// uniform brick of ranks over the box (procmap.cpp / Comm::set_proc_grid): unless `processors` fixed it,
// all ranks form slabs along the longest periodic axis (else the longest axis) -- NVSwitch gives every
// pair of GPUs the same bandwidth, so 1-D slabs minimise the number of messages per step
static void setup_decomposition(dem_engine *E)
{
  if (!E->user_grid) {
    int best = 0; double bl = -1.0;
    for (int d = 0; d < 3; d++) { const double l = E->prd[d] * (E->periodic[d] ? 1.0 : 0.999); if (l > bl) { bl = l; best = d; } }
    E->pgrid[0] = E->pgrid[1] = E->pgrid[2] = 1; E->pgrid[best] = E->nranks;
  }
  int r = E->rank;
  E->myloc[0] = r % E->pgrid[0]; r /= E->pgrid[0];
  E->myloc[1] = r % E->pgrid[1]; r /= E->pgrid[1];
  E->myloc[2] = r;
  for (int d = 0; d < 3; d++) {
    E->sublo[d] = E->lo[d] + E->prd[d] * E->myloc[d] / E->pgrid[d];
    E->subhi[d] = (E->myloc[d] == E->pgrid[d] - 1) ? E->hi[d] : E->lo[d] + E->prd[d] * (E->myloc[d] + 1) / E->pgrid[d];
  }
}
static int rank_of(dem_engine *E, int ix, int iy, int iz) { return (iz * E->pgrid[1] + iy) * E->pgrid[0] + ix; }
static int neighbor_rank(dem_engine *E, int dim, int dir)
{  // dir -1 / +1; returns -1 when there is no neighbour across a non-periodic face
  int loc[3] = {E->myloc[0], E->myloc[1], E->myloc[2]};
  loc[dim] += dir;
  if (loc[dim] < 0) { if (!E->periodic[dim]) return -1; loc[dim] = E->pgrid[dim] - 1; }
  if (loc[dim] >= E->pgrid[dim]) { if (!E->periodic[dim]) return -1; loc[dim] = 0; }
  return rank_of(E, loc[0], loc[1], loc[2]);
}

extern "C" int dem_decomposition(dem_engine *e, int pgrid[3], int myloc[3], double sublo[3], double subhi[3])
{
  API_BEGIN
  setup_decomposition(e);
  for (int d = 0; d < 3; d++) { if (pgrid) pgrid[d] = e->pgrid[d]; if (myloc) myloc[d] = e->myloc[d]; if (sublo) sublo[d] = e->sublo[d]; if (subhi) subhi[d] = e->subhi[d]; }
  API_END
}

static MineP brick_params(const dem_engine *e)
{
  MineP B;
  for (int d = 0; d < 3; d++) {
    B.lo[d] = e->lo[d]; B.hi[d] = e->hi[d]; B.prd[d] = e->prd[d]; B.sublo[d] = e->sublo[d]; B.subhi[d] = e->subhi[d];
    B.periodic[d] = e->periodic[d]; B.first[d] = e->myloc[d] == 0; B.last[d] = e->myloc[d] == e->pgrid[d] - 1;
  }
  return B;
}
// Decomposition without a device: the brick of `rank` among `nranks` over a box (processors grid optional: px*py*pz == nranks,
// or null / zeros for the engine's own choice), its six face neighbours (-1 = none) and, for `n` positions, whether this rank
// owns them -- the very functions dem_upload_particles and the halo code use (procmap.cpp / comm_brick.cpp:215-330 analogue).
// Needs no GPU: lets multi-process host logic be tested with any torch.distributed backend.
extern "C" int dem_brick_layout(int nranks, int rank, const double lo[3], const double hi[3], const int periodic[3], const int *procgrid,
                                int pgrid[3], int myloc[3], double sublo[3], double subhi[3], int neigh[6],
                                long n, const double *x, int *mine)
{
  if (nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi || !periodic) return DEM_ERR_ARG;
  dem_engine E;
  E.nranks = nranks; E.rank = rank;
  for (int d = 0; d < 3; d++) { E.lo[d] = lo[d]; E.hi[d] = hi[d]; E.prd[d] = hi[d] - lo[d]; E.periodic[d] = periodic[d]; if (!(hi[d] > lo[d])) return DEM_ERR_ARG; }
  if (procgrid && procgrid[0] * procgrid[1] * procgrid[2] != 0) {
    if (procgrid[0] * procgrid[1] * procgrid[2] != nranks) return DEM_ERR_ARG;
    for (int d = 0; d < 3; d++) E.pgrid[d] = procgrid[d];
    E.user_grid = 1;
  }
  setup_decomposition(&E);
  for (int d = 0; d < 3; d++) {
    if (pgrid) pgrid[d] = E.pgrid[d]; if (myloc) myloc[d] = E.myloc[d]; if (sublo) sublo[d] = E.sublo[d]; if (subhi) subhi[d] = E.subhi[d];
    if (neigh) { neigh[2 * d] = E.pgrid[d] > 1 ? neighbor_rank(&E, d, -1) : -1; neigh[2 * d + 1] = E.pgrid[d] > 1 ? neighbor_rank(&E, d, 1) : -1; }
  }
  if (n > 0 && x && mine) { const MineP B = brick_params(&E); for (long i = 0; i < n; i++) mine[i] = brick_owns(B, x + 3 * i) ? 1 : 0; }
  return DEM_OK;
}

static void ensure_cub(dem_engine *E, size_t n);
static int compact_flags(dem_engine *E, int n, DevBuf<int> &flag, DevBuf<int> &scan, DevBuf<int> &list);
extern "C" int dem_upload_particles(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x,
                                    const double *v, const double *omega, const double *radius, const double *density)
{
  API_BEGIN
  if (n < 0 || (n > 0 && (!tag || !type || !x || !radius || !density))) dem_fail(e, DEM_ERR_ARG, "missing particle arrays");
  if (n >= (long)NBR_IDX) dem_fail(e, DEM_ERR_OVERFLOW, "more than 2^25 particles on one GPU (neighbour words index owned + ghost particles in 25 bits)");
  CK(cudaSetDevice(e->device));
  e->cap = 0;  // force fresh allocation
  for (int b = 0; b < 2; b++) { e->xr[b].release(); e->vm[b].release(); e->wt[b].release(); }
  e->xh.release(); e->tag.release(); e->density.release(); e->f.release(); e->tq.release(); e->whist.release();
  setup_decomposition(e);
  if (e->nranks > 1 && n > 0 && !getenv("DEM_B200_HOST_UPLOAD")) {
    // several bricks: every rank ships the caller's arrays as they are; the brick's particles are selected, compacted (in
    // input order) and packed into records on the device
    cudaStream_t st = e->stream;
    const size_t nd = (size_t)n;
    const size_t bytes = nd * (3 + 3 + 3 + 1 + 1) * sizeof(double) + nd * 3 * sizeof(int) + 64;
    e->stage.ensure(e, bytes);
    double *dx = (double *)e->stage.p, *dv = dx + 3 * nd, *dw = dv + 3 * nd, *dr = dw + 3 * nd, *dd = dr + nd;
    int *dt = (int *)(dd + nd), *dm = dt + nd, *dg = dm + nd;
    CK(cudaMemcpyAsync(dx, x, 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    if (v) CK(cudaMemcpyAsync(dv, v, 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    if (omega) CK(cudaMemcpyAsync(dw, omega, 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dr, radius, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dd, density, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dt, type, nd * sizeof(int), cudaMemcpyHostToDevice, st));
    if (mask) CK(cudaMemcpyAsync(dm, mask, nd * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dg, tag, nd * sizeof(int), cudaMemcpyHostToDevice, st));
    e->counters.ensure(e, 4);
    CK(cudaMemsetAsync(e->counters.p, 0, 2 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(e->counters.p + 2, 0xFF, sizeof(unsigned long long), st));  // running minimum of the radius bits
    const MineP B = brick_params(e);
    e->flo.ensure(e, nd + 1); e->slo.ensure(e, nd + 1);
    ensure_cub(e, nd);
    k_flag_mine<<<GRID(n, 256), 256, 0, st>>>((int)n, dx, dr, dd, dt, dg, e->ntypes, B, e->flo.p, (int *)(e->counters.p + 1), e->counters.p);
    e->launches++;
    DevBuf<int> list;
    const int nmine = compact_flags(e, (int)n, e->flo, e->slo, list);
    const double cf = e->opt.count("cap_factor") ? e->opt["cap_factor"] : 1.5;
    ensure_particle_cap(e, std::max<long>((long)(nmine * cf) + 1024, 1024), 0);
    e->cur = 0;
    if (nmine) {
      k_pack_upload_sel<<<GRID(nmine, 256), 256, 0, st>>>(nmine, list.p, dx, v ? dv : nullptr, omega ? dw : nullptr, dr, dd, dt, mask ? dm : nullptr, dg,
                                                          e->xr[0].p, e->vm[0].p, e->wt[0].p, e->tag.p, e->density.p, e->ins_mass);
      e->launches++;
    }
    CK(cudaMemsetAsync(e->xh.p, 0, (size_t)e->cap * sizeof(double4), st));
    unsigned long long hc[3];
    CK(cudaMemcpyAsync(hc, e->counters.p, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    list.release();
    const int errbits = (int)(hc[1] & 0xffffffffu);
    if (errbits & 1) dem_fail(e, DEM_ERR_ARG, "Invalid atom type in particle data");
    if (errbits & 2) dem_fail(e, DEM_ERR_ARG, "Invalid radius or density in particle data");
    if (errbits & 4) dem_fail(e, DEM_ERR_ARG, "Invalid atom ID in particle data");
    double rmaxd; memcpy(&rmaxd, &hc[0], 8);
    { double rm; memcpy(&rm, &hc[2], 8); e->rmin = rm; }  // (minimum over the uploaded set, found on the device)
    e->nlocal = nmine; e->nghost = 0; e->rmax = rmaxd;
    e->uploaded = 1; e->setup_done = 0; e->forces_valid = 0; e->order_valid = 0; e->mesh_ready = 0;
    e->ls[0].valid = e->ls[1].valid = 0;
    e->ntimestep = 0;
    return DEM_OK;
  }
  if (e->nranks == 1 && n > 0) {
    // single brick: ship the caller's arrays as they are and build the records on the device
    const double cf = e->opt.count("cap_factor") ? e->opt["cap_factor"] : 1.25;
    ensure_particle_cap(e, std::max<long>((long)(n * cf) + 1024, 1024), 0);
    cudaStream_t st = e->stream;
    const size_t nd = (size_t)n;
    const size_t bytes = nd * (3 + 3 + 3 + 1 + 1) * sizeof(double) + nd * 3 * sizeof(int) + 64;
    e->stage.ensure(e, bytes);
    double *dx = (double *)e->stage.p, *dv = dx + 3 * nd, *dw = dv + 3 * nd, *dr = dw + 3 * nd, *dd = dr + nd;
    int *dt = (int *)(dd + nd), *dm = dt + nd, *dg = dm + nd;
    CK(cudaMemcpyAsync(dx, x, 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    if (v) CK(cudaMemcpyAsync(dv, v, 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    if (omega) CK(cudaMemcpyAsync(dw, omega, 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dr, radius, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dd, density, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dt, type, nd * sizeof(int), cudaMemcpyHostToDevice, st));
    if (mask) CK(cudaMemcpyAsync(dm, mask, nd * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dg, tag, nd * sizeof(int), cudaMemcpyHostToDevice, st));
    e->counters.ensure(e, 4);
    CK(cudaMemsetAsync(e->counters.p, 0, 2 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(e->counters.p + 2, 0xFF, sizeof(unsigned long long), st));  // running minimum of the radius bits
    e->cur = 0;
    k_pack_upload<<<GRID(n, 256), 256, 0, st>>>((int)n, dx, v ? dv : nullptr, omega ? dw : nullptr, dr, dd, dt, mask ? dm : nullptr, dg, e->ntypes,
                                                 e->xr[0].p, e->vm[0].p, e->wt[0].p, (int *)(e->counters.p + 1), e->counters.p, e->ins_mass);
    e->launches++;
    CK(cudaMemcpyAsync(e->tag.p, dg, nd * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(e->density.p, dd, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(e->xh.p, 0, (size_t)e->cap * sizeof(double4), st));
    unsigned long long hc[3];
    CK(cudaMemcpyAsync(hc, e->counters.p, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int errbits = (int)(hc[1] & 0xffffffffu);
    if (errbits & 1) dem_fail(e, DEM_ERR_ARG, "Invalid atom type in particle data");
    if (errbits & 2) dem_fail(e, DEM_ERR_ARG, "Invalid radius or density in particle data");
    if (errbits & 4) dem_fail(e, DEM_ERR_ARG, "Invalid atom ID in particle data");
    double rmaxd; memcpy(&rmaxd, &hc[0], 8);
    { double rm; memcpy(&rm, &hc[2], 8); e->rmin = rm; }  // (minimum over the uploaded set, found on the device)
    e->nlocal = n; e->nghost = 0; e->rmax = rmaxd;
    e->uploaded = 1; e->setup_done = 0; e->forces_valid = 0; e->order_valid = 0; e->mesh_ready = 0;
    e->ls[0].valid = e->ls[1].valid = 0;
    e->ntimestep = 0;
    return DEM_OK;
  }
  std::vector<double4> hx, hv, hw; std::vector<int> htag; std::vector<double> hden;
  hx.reserve(n / e->nranks + 16); hv.reserve(n / e->nranks + 16); hw.reserve(n / e->nranks + 16);
  double rmax = 0.0;
  for (long i = 0; i < n; i++) {
    if (type[i] < 1 || type[i] > e->ntypes) dem_fail(e, DEM_ERR_ARG, "Invalid atom type in particle %ld", i);
    if (!(radius[i] > 0.0) || !(density[i] > 0.0)) dem_fail(e, DEM_ERR_ARG, "Invalid radius or density in particle %ld", i);
    if (tag[i] <= 0) dem_fail(e, DEM_ERR_ARG, "Invalid atom ID in particle %ld", i);
    const double r = radius[i];
    const double m = 4.0 * 3.14159265358979323846 / 3.0 * r * r * r * density[i];  // atom_vec_sphere.cpp:1078
    rmax = std::max(rmax, r);  // global maximum: every rank sees the full set
    e->rmin = (i == 0) ? r : std::min(e->rmin, r);
    if (e->nranks > 1 && !brick_owns(brick_params(e), x + 3 * i)) continue;  // ownership on the wrapped position (Domain::pbc + sub-box, like read_data)
    hx.push_back(make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], r));
    hv.push_back(make_double4(v ? v[3 * i] : 0., v ? v[3 * i + 1] : 0., v ? v[3 * i + 2] : 0., m));
    const long long bits = pack_bits(type[i], mask ? mask[i] : 1);
    double wb; memcpy(&wb, &bits, 8);
    hw.push_back(make_double4(omega ? omega[3 * i] : 0., omega ? omega[3 * i + 1] : 0., omega ? omega[3 * i + 2] : 0., wb));
    htag.push_back(tag[i]); hden.push_back(density[i]);
  }
  n = (long)hx.size();
  tag = htag.data(); density = hden.data();
  {
    const double cf = e->opt.count("cap_factor") ? e->opt["cap_factor"] : (e->nranks > 1 ? 1.5 : 1.25);
    ensure_particle_cap(e, std::max<long>((long)(n * cf) + 1024, 1024), 0);
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
  e->uploaded = 1; e->setup_done = 0; e->forces_valid = 0; e->order_valid = 0; e->mesh_ready = 0;
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
  bool hertz = false, hooke = false, tang = false, roll = false, epsd = false, hyst = false;
  auto scan = [&](const ModelP &m) { hertz |= m.normal == N_HERTZ; hooke |= m.normal == N_HOOKE; hyst |= m.normal == N_HYST1 || m.normal == N_HYST2; tang |= m.tangential != 0; roll |= m.rolling != R_OFF; epsd |= m.rolling == R_EPSD; };
  if (E->have_pair) scan(E->pm);
  for (auto &w : E->walls) scan(w.p.m);
  for (auto &w : E->mwalls) scan(w.m);
  if (hertz || hooke) { need("youngsModulus"); need("poissonsRatio"); need("coefficientRestitution"); }
  if (hooke) need("characteristicVelocity");
  if (hyst) for (const char *k : {"coefficientRestitution", "LoadingStiffness", "UnloadingStiffness", "coefficientAdhesionStiffness", "coefficientPlasticityDepth", "pullOffForce",
                                  "alphaCustom", "cinCustom", "aoneCustom", "atwoCustom", "athreeCustom", "kcinCustom"}) need(k);
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
    if (hyst) {
      if (!(hertz || hooke)) { const double cr = E->cor[i][j]; if (cr <= 0.05 || cr > 1) dem_fail(E, DEM_ERR_ARG, "0.05 < coefficientRestitution <= 1 required"); at(T_CORLOG) = log(cr); }
      for (int w = T_H_KEL; w <= T_H_KCIN; w++) at(w) = E->bp[w][i][j];
    }
    at(T_MU) = E->mu[i][j]; at(T_RMU) = E->rmu[i][j]; at(T_RVISC) = E->rvisc[i][j];
    if (hertz || hooke) { at(T_SQ2Y) = sqrt(2. * at(T_YEFF)); at(T_SQ8G) = sqrt(8. * at(T_GEFF)); at(T_INV8G) = 1. / (8. * at(T_GEFF)); }
  }
  E->cdf = E->cdf_user > 1.0 ? E->cdf_user : 1.0;  // neigh_modify contact_distance_factor (neighbor.cpp:1922-1925)
  if (E->cdf > 1.0 && E->have_pair && E->pm.normal >= N_HYST1) dem_fail(E, DEM_ERR_UNSUPPORTED, "contact_distance_factor > 1 with the hysteretic/nonlinear laws is outside the hot-path scope");
  if (E->have_pair && E->pm.cohesion) {
    const ModelP &m = E->pm;
    const bool nl = m.cohesion == C_BONDNL;
    const char *sfx = nl ? "nonlinear" : "";
    auto needb = [&](const std::string &base) { const std::string nm = base + sfx; need(nm.c_str()); };
    needb("radiusMultiplierBond"); needb("createDistanceBond");
    if (!m.createAlways) needb("tsCreateBond");
    if (!m.damping && !m.dissipation) dem_fail(E, DEM_ERR_ARG, "Damping or dissipation has to be enabled.");
    if (m.dissipation) for (const char *k : {"dissipationNormalForceBond", "dissipationTangentialForceBond", "dissipationNormalTorqueBond", "dissipationTangentialTorqueBond"}) need(k);
    if (m.damping) { needb("dampingNormalForceBond"); needb("dampingTangentialForceBond"); needb("dampingNormalTorqueBond"); needb("dampingTangentialTorqueBond"); }
    if (!m.stressBreak) needb("maxDistanceBond"); else { needb("maxSigmaBond"); needb("maxTauBond"); }
    if (!nl) { need("normalBondStiffnessPerUnitArea"); need("tangentialBondStiffnessPerUnitArea"); }
    else for (const char *k : {"K_fn1", "Ku_fn1", "Kc_fn1", "K_fn2", "Ku_fn2", "Kc_fn2", "K_ft", "K_tn", "Ku_tn", "Kc_tn", "K_tt", "Ku_tt", "Kc_tt"}) need((std::string("stiffnessPerUnitArea") + k).c_str());
    for (int w = T_B_LAMBDA; w < T_H_KEL; w++) for (int i = 1; i <= T; i++) for (int j = 1; j <= T; j++) t[((size_t)w * n1 + i) * n1 + j] = E->bp[w][i][j];
    if (m.dissipation) for (int w = T_D_FN; w <= T_D_TT; w++) for (int i = 1; i <= T; i++) for (int j = 1; j <= T; j++) {
      if (E->bp[w][i][j] < E->dt) dem_fail(E, DEM_ERR_ARG, "dissipation time scale > time-step size required");
      t[((size_t)w * n1 + i) * n1 + j] = 1. / E->bp[w][i][j];  // saved already inverted
    }
    // neighbor->register_contact_dist_factor: cohesion_model_bond.h:391-475, cohesion_model_bond_nonlinear.h:336-382
    if (!(E->rmin > 0.)) dem_fail(E, DEM_ERR_STATE, "Bond settings: The minimum radius can't be <= 0!");
    double cdf_all = E->cdf;  // register_contact_dist_factor keeps the larger one (neighbor.h:142)
    for (int i = 1; i <= T; i++) for (int j = 1; j <= T; j++) {
      double one;
      if (!m.stressBreak) one = 1.1 * 0.5 * E->bp[T_B_MAXDIST][i][j] / E->rmin;
      else {
        double stress = E->bp[T_B_MAXSIGMA][i][j];
        if (m.ratioTC) stress = fmax(stress, stress * E->bp[T_B_RATIOTC][i][j]);
        const double kk = nl ? E->bp[T_B_K_FN2][i][j] : E->bp[T_B_KN][i][j];
        if (kk <= 1e-15) dem_fail(E, DEM_ERR_ARG, "Bond settings: In case of stress breakage, the normal bond stiffness can't be <= 0!");
        one = nl ? 0.5 * 1.1 * E->bp[T_B_CREATEDIST][i][j] / E->rmin + 0.5 * 1.1 * stress / (kk * E->rmin)
                 : 0.5 * (1.1 * E->bp[T_B_CREATEDIST][i][j] / E->rmin + 1.1 * stress / (kk * E->rmin));
      }
      cdf_all = one > cdf_all ? one : cdf_all;
    }
    if (cdf_all > 10.) dem_fail(E, DEM_ERR_ARG, "Maximum bond distance exceeding 10 x particle diameter, please reduce maxDistanceBond or maxSigmaBond/maxTauBond");
    E->cdf = cdf_all;
  }
  for (int w = 0; w < T_B_LAMBDA; w++) E->t1[w] = t[((size_t)w * n1 + 1) * n1 + 1];
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
{  // cell grid over this rank's sub-box plus one layer of cells for the ghosts (neighbor.cpp:1647-1812 analogue)
  E->cutneighmax = 2.0 * E->rmax * E->cdf + E->skin;  // pair_gran.cpp:591-603 + neighbor skin
  if (!(E->cutneighmax > 0)) dem_fail(E, DEM_ERR_STATE, "neighbour cutoff is zero (no particles or zero radius)");
  long total = 1;
  double cell = E->cutneighmax;
  for (int pass = 0; pass < 64; pass++) {
    total = 1;
    for (int d = 0; d < 3; d++) {
      const double len = E->subhi[d] - E->sublo[d];
      int nc = (int)floor(len / cell); if (nc < 1) nc = 1;
      const double size = len / nc;
      E->grid.nc[d] = nc + 2; E->grid.inv[d] = 1.0 / size; E->grid.org[d] = E->sublo[d] - size;
      total *= (nc + 2);
    }
    if (total <= (1L << 25)) break;
    cell *= 1.3;
  }
  for (int d = 0; d < 3; d++) {
    if (E->periodic[d] && E->prd[d] < 2.0 * E->cutneighmax)
      dem_fail(E, DEM_ERR_UNSUPPORTED, "periodic box length in dim %d is below two neighbour cutoffs", d);
    if (E->pgrid[d] > 1 && E->subhi[d] - E->sublo[d] < 2.0 * E->cutneighmax)
      dem_fail(E, DEM_ERR_UNSUPPORTED, "sub-box of a rank in dim %d is thinner than two neighbour cutoffs", d);
  }
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

// exchange `n` ints with the two peers of a swap (either may be -1)
static void xchg_ints(dem_engine *E, int peer_send, const int *sendv, int peer_recv, int *recvv, int n)
{
  cudaStream_t st = E->stream;
  E->cnt_dev.ensure(E, 64);
  for (int k = 0; k < n; k++) { E->hcnt[k] = sendv[k]; E->hcnt[32 + k] = 0; }
  CK(cudaMemcpyAsync(E->cnt_dev.p, E->hcnt, n * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(E->cnt_dev.p + 32, 0, n * sizeof(int), st));
  NK(g_nccl.GroupStart());
  if (peer_send >= 0) NK(g_nccl.Send(E->cnt_dev.p, n, ncclInt, peer_send, E->comm, st));
  if (peer_recv >= 0) NK(g_nccl.Recv(E->cnt_dev.p + 32, n, ncclInt, peer_recv, E->comm, st));
  NK(g_nccl.GroupEnd());
  CK(cudaMemcpyAsync(E->hcnt + 32, E->cnt_dev.p + 32, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int k = 0; k < n; k++) recvv[k] = E->hcnt[32 + k];
}

// blocking exchange of a small byte block with the two peers of a swap (either may be -1)
static void xchg_bytes(dem_engine *E, int peer_send, const void *sendv, int peer_recv, void *recvv, size_t nbytes)
{
  cudaStream_t st = E->stream;
  E->stage.ensure(E, 2 * nbytes + 64);
  char *d = E->stage.p;
  CK(cudaMemcpyAsync(d, sendv, nbytes, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d + nbytes, 0, nbytes, st));
  NK(g_nccl.GroupStart());
  if (peer_send >= 0) NK(g_nccl.Send(d, nbytes, ncclChar, peer_send, E->comm, st));
  if (peer_recv >= 0) NK(g_nccl.Recv(d + nbytes, nbytes, ncclChar, peer_recv, E->comm, st));
  NK(g_nccl.GroupEnd());
  CK(cudaMemcpyAsync(recvv, d + nbytes, nbytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
}

// End of a rebuild: every rank tells the rank that will push ghosts to it where they go (CUDA IPC handles of its six
// record arrays and of its signal words, the first ghost index of the swap, its buffer parity).  From here until the
// next rebuild the per-step ghost refresh is plain NVLink stores issued by the sender's pack kernel -- no NCCL call.
struct HaloInfo { cudaIpcMemHandle_t h[7]; int gfirst, cur, ok, gen; };
// CUDA-IPC mappings of the other ranks' blocks live in a process-wide cache keyed by the handle bytes: the ranks recycle
// their device blocks between engines (dev_alloc), so the next engine of a job meets the very same handles and must not pay
// cudaIpcOpenMemHandle again (measured: 44 ms of a 53 ms end-to-end job at 8 ranks went to re-opening 21 handles).  A mapping
// stays valid as long as the exporting process keeps the block, i.e. until its dem_trim_memory; ours are closed there too.
struct IpcKey { unsigned char b[64]; bool operator<(const IpcKey &o) const { return memcmp(b, o.b, 64) < 0; } };
static std::map<std::pair<int, IpcKey>, void *> g_ipc;  // (device, handle) -> mapping
static void ipc_close_all()
{
  std::lock_guard<std::mutex> lk(g_mem_mu);
  int cur = 0; cudaGetDevice(&cur);
  for (auto &kv : g_ipc) { cudaSetDevice(kv.first.first); cudaIpcCloseMemHandle(kv.second); }
  g_ipc.clear();
  cudaSetDevice(cur);
}
static void *ipc_open(dem_engine *E, const cudaIpcMemHandle_t &h, int rank, int gen)
{
  (void)rank; (void)gen;
  IpcKey k; memcpy(k.b, &h, 64);
  std::lock_guard<std::mutex> lk(g_mem_mu);
  auto it = g_ipc.find({E->device, k});
  if (it != g_ipc.end()) return it->second;
  void *p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_ipc[{E->device, k}] = p;
  return p;
}
static void halo_p2p_setup(dem_engine *E)
{
  if (E->nranks == 1) return;
  const bool want = !(E->opt.count("p2p") && E->opt["p2p"] == 0);
  cudaStream_t st = E->stream;
  E->hsig.ensure(E, 16);
  CK(cudaMemsetAsync(E->hsig.p, 0, 16 * sizeof(int), st));
  CK(cudaStreamSynchronize(st));
  E->cur0 = E->cur;
  int all_ok = 1;
  for (int q = 0; q < E->nswap; q++) {
    dem_engine::Swap &W = E->swaps[q];
    W.p2p = 0; W.serial = 0;
    if (W.self) continue;
    const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1), peer_recv = neighbor_rank(E, W.dim, W.side ? -1 : 1);
    HaloInfo mine, theirs;
    memset(&mine, 0, sizeof mine); memset(&theirs, 0, sizeof theirs);
    mine.ok = want ? 1 : 0;
    if (want) {
      void *bufs[7] = {E->xr[0].p, E->xr[1].p, E->vm[0].p, E->vm[1].p, E->wt[0].p, E->wt[1].p, E->hsig.p};
      for (int k = 0; k < 7; k++) if (cudaIpcGetMemHandle(&mine.h[k], bufs[k]) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    }
    mine.gfirst = W.gfirst; mine.cur = E->cur; mine.gen = E->alloc_gen;
    // my info goes to the rank that sends to me in this swap; I receive the info of the rank I send to
    xchg_bytes(E, peer_recv, &mine, peer_send, &theirs, sizeof(HaloInfo));
    if (peer_send < 0) continue;
    int ok = theirs.ok && want;
    if (ok) {
      for (int k = 0; k < 6 && ok; k++) { W.pbase[k] = (double4 *)ipc_open(E, theirs.h[k], peer_send, theirs.gen); if (!W.pbase[k]) ok = 0; }
      W.psig = ok ? (int *)ipc_open(E, theirs.h[6], peer_send, theirs.gen) : nullptr;
      if (!W.psig) ok = 0;
    }
    W.pgfirst = theirs.gfirst; W.pcur0 = theirs.cur; W.p2p = ok;
    if (!ok) all_ok = 0;
  }
  // every rank must take the same path for a given swap pair: agree globally
  E->cnt_dev.ensure(E, 64);
  E->hcnt[0] = all_ok;
  CK(cudaMemcpyAsync(E->cnt_dev.p, E->hcnt, sizeof(int), cudaMemcpyHostToDevice, st));
  NK(g_nccl.AllReduce(E->cnt_dev.p, E->cnt_dev.p + 1, 1, ncclInt, ncclMin, E->comm, st));
  CK(cudaMemcpyAsync(E->hcnt + 1, E->cnt_dev.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  E->p2p_ok = E->hcnt[1];
  if (!E->p2p_ok) for (int q = 0; q < E->nswap; q++) E->swaps[q].p2p = 0;
  if (E->p2p_ok && !E->fbox_ready && E->nranks <= DEM_MAXRANKS && !(E->opt.count("peer_flags") && E->opt["peer_flags"] == 0)) {
    // flag boxes: every rank maps every rank's box once (the box is never reallocated).  All ranks take the same decision.
    E->fbox.ensure(E, FBOX_INTS);
    CK(cudaMemsetAsync(E->fbox.p, 0, FBOX_INTS * sizeof(int), st));
    cudaIpcMemHandle_t mine; int ok = cudaIpcGetMemHandle(&mine, E->fbox.p) == cudaSuccess ? 1 : 0;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof mine); }
    E->stage.ensure(E, 64 * (size_t)(E->nranks + 1) + 64);
    CK(cudaMemcpyAsync(E->stage.p, &mine, 64, cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllGather(E->stage.p, E->stage.p + 64, 64, ncclChar, E->comm, st));
    std::vector<cudaIpcMemHandle_t> all(E->nranks);
    CK(cudaMemcpyAsync(all.data(), E->stage.p + 64, 64 * (size_t)E->nranks, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int r = 0; r < E->nranks && ok; r++) {
      if (r == E->rank) { E->peer_fbox[r] = E->fbox.p; continue; }
      E->peer_fbox[r] = (int *)ipc_open(E, all[r], -1 - r, 0);  // (pseudo rank: the record arrays' generation sweep never closes it)
      if (!E->peer_fbox[r]) ok = 0;
    }
    E->hcnt[0] = ok;
    CK(cudaMemcpyAsync(E->cnt_dev.p, E->hcnt, sizeof(int), cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllReduce(E->cnt_dev.p, E->cnt_dev.p + 1, 1, ncclInt, ncclMin, E->comm, st));
    CK(cudaMemcpyAsync(E->hcnt + 1, E->cnt_dev.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    E->fbox_ready = E->hcnt[1] ? 1 : -1;  // -1: tried, not available -> NCCL all-reduce per step
    if (E->fbox_ready == 1 && !E->hflag_dev) { void *dp = nullptr; if (cudaHostGetDevicePointer(&dp, E->hflag, 0) != cudaSuccess) { cudaGetLastError(); E->fbox_ready = -1; } E->hflag_dev = (int *)dp; }
  }
}

// developer aid: DEM_B200_TRACE=1 prints the host+device time of each rebuild stage (synchronises at every mark)
struct StageTrace {
  bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0;
  explicit StageTrace(cudaStream_t s) : on(getenv("DEM_B200_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what)
  {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[dem trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// ---- fused ghost push.  After the borders of a rebuild every ghost slot of every rank is a copy of ONE owned particle of
// this rank or of a neighbour rank (one decomposed dimension: a ghost crosses at most one rank boundary; the periodic images
// of the other dimensions are made on the rank).  Here every rank learns, per owned particle, all the slots that hold a copy
// of it -- "direct" ghosts on the two neighbours (entry k of my send list = their slot pgfirst + k), the periodic images the
// neighbours make of those ghosts (they report them: swap, entry, slot, shift), and its own periodic images -- so that the
// step kernel can store a particle's new records into all of them from its epilogue (step_epilogue); one k_handshake launch
// per step then publishes the exchange and the step flags and waits for the neighbours' (k_flags_host on a single rank).
// comm_brick.cpp:563-645 (forward_comm) is what this replaces between two rebuilds: no pack kernel, no send/recv, no wait
// launch per swap.  The table is assembled on the device (k_img_chain / k_img_local / k_img_remote, linked entries per
// particle); any unsupported constellation keeps the k_push path.
static void fused_halo_setup(dem_engine *E)
{
  E->fz_ready = 0;
  if (E->opt.count("fused_halo") && E->opt["fused_halo"] == 0) return;
  const bool solo = E->nranks == 1;  // one rank: only its own periodic images (fz_ready = 2)
  if (!solo && (!E->p2p_ok || E->fbox_ready != 1)) return;
  int ndec = 0;
  for (int d = 0; d < 3; d++) if (E->pgrid[d] > 1) ndec++;
  if (!solo && ndec != 1) return;
  int rq[2] = {-1, -1};
  for (int q = 0; q < E->nswap; q++) if (!E->swaps[q].self) rq[E->swaps[q].side] = q;
  if (!solo && (rq[0] < 0 || rq[1] < 0)) return;
  if (solo && E->nswap == 0) return;
  cudaStream_t st = E->stream;
  const int nl = (int)E->nlocal, ng = (int)E->nghost;
  auto shiftcode = [](const dem_engine::Swap &W) { return W.shift == 0.0 ? 0 : ((W.shift > 0.0 ? 1 : 2) << (2 * W.dim)); };
  // work space: gsrc [ng] int4, two report buffers [3 ng] int, counters [4]
  E->img_ws.ensure(E, (size_t)4 * std::max(ng, 1) + (size_t)6 * std::max(ng, 1) + 16);
  int4 *gsrc = (int4 *)E->img_ws.p;
  int *rep[2] = {E->img_ws.p + 4 * (size_t)std::max(ng, 1), E->img_ws.p + 7 * (size_t)std::max(ng, 1)};
  int *cnt = E->img_ws.p + 10 * (size_t)std::max(ng, 1);
  CK(cudaMemsetAsync(cnt, 0, 4 * sizeof(int), st));
  for (int q = 0; q < E->nswap; q++) {  // origin of every ghost slot, swap after swap
    const dem_engine::Swap &W = E->swaps[q];
    if (!W.nrecv) continue;
    if (W.gfirst < nl || W.gfirst + W.nrecv > nl + ng) return;
    k_img_chain<<<GRID(W.nrecv, 256), 256, 0, st>>>(W.nrecv, W.list.p, nl, W.gfirst, W.self, W.side, shiftcode(W), gsrc, rep[0], rep[1], cnt);
    E->launches++;
  }
  int hc[4] = {0, 0, 0, 0};
  if (!solo) { CK(cudaMemcpyAsync(hc, cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
  // the neighbours' reports about the ghosts I sent them
  int nrep_in[2] = {0, 0};
  for (int sd = 0; sd < 2 && !solo; sd++) {
    const dem_engine::Swap &W = E->swaps[rq[sd]];
    const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1), peer_recv = neighbor_rank(E, W.dim, W.side ? -1 : 1);
    // my reports about this side's swap go to the rank that sent it to me; I get the reports of the rank I sent it to
    int sv[1] = {hc[sd]}, rv[1] = {0};
    xchg_ints(E, peer_recv, sv, peer_send, rv, 1);
    nrep_in[sd] = rv[0];
  }
  const size_t ntab = (size_t)ng + (solo ? 0 : (size_t)E->swaps[rq[0]].nsend + E->swaps[rq[1]].nsend + nrep_in[0] + nrep_in[1]);
  E->img_first.ensure(E, (size_t)std::max(nl, 1)); E->img_tab.ensure(E, ntab + 1); E->imgp.ensure(E, 1);
  E->img_in.ensure(E, 3 * (size_t)(nrep_in[0] + nrep_in[1]) + 4);
  CK(cudaMemsetAsync(E->img_first.p, 0xff, sizeof(int) * (size_t)std::max(nl, 1), st));
  if (ng) { k_img_local<<<GRID(ng, 256), 256, 0, st>>>(ng, nl, gsrc, E->img_first.p, E->img_tab.p); E->launches++; }
  size_t base = ng;
  for (int sd = 0; sd < 2 && !solo; sd++) {
    dem_engine::Swap &W = E->swaps[rq[sd]];
    const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1), peer_recv = neighbor_rank(E, W.dim, W.side ? -1 : 1);
    int *in = E->img_in.p + 3 * (size_t)(sd ? nrep_in[0] : 0);
    if ((peer_recv >= 0 && hc[sd]) || (peer_send >= 0 && nrep_in[sd])) {
      NK(g_nccl.GroupStart());
      if (peer_recv >= 0 && hc[sd]) NK(g_nccl.Send(rep[sd], 3 * (size_t)hc[sd], ncclInt, peer_recv, E->comm, st));
      if (peer_send >= 0 && nrep_in[sd]) NK(g_nccl.Recv(in, 3 * (size_t)nrep_in[sd], ncclInt, peer_send, E->comm, st));
      NK(g_nccl.GroupEnd());
    }
    if (peer_send >= 0 && W.p2p) {
      const int sc = shiftcode(W);
      if (W.nsend) k_img_remote<<<GRID(W.nsend, 256), 256, 0, st>>>(W.nsend, 1, nullptr, W.list.p, W.nsend, nl, gsrc, 1 + sd, W.pgfirst, sc, (int)base, E->img_first.p, E->img_tab.p, cnt);
      if (nrep_in[sd]) k_img_remote<<<GRID(nrep_in[sd], 256), 256, 0, st>>>(nrep_in[sd], 0, in, W.list.p, W.nsend, nl, gsrc, 1 + sd, W.pgfirst, sc, (int)(base + W.nsend), E->img_first.p, E->img_tab.p, cnt);
      E->launches += 2;
    }
    base += (size_t)W.nsend + nrep_in[sd];
  }
  ImgP I; memset(&I, 0, sizeof I);
  for (int b = 0; b < 2; b++) { I.bx[0][b] = E->xr[b].p; I.bv[0][b] = E->vm[b].p; I.bw[0][b] = E->wt[b].p; }
  if (solo) E->cur0 = E->cur;
  I.pcur0[0] = E->cur0;
  for (int sd = 0; sd < 2 && !solo; sd++) {
    dem_engine::Swap &W = E->swaps[rq[sd]];
    const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1);
    if (peer_send < 0 || !W.p2p) continue;
    for (int b = 0; b < 2; b++) { I.bx[1 + sd][b] = W.pbase[0 + b]; I.bv[1 + sd][b] = W.pbase[2 + b]; I.bw[1 + sd][b] = W.pbase[4 + b]; }
    I.pcur0[1 + sd] = W.pcur0;
  }
  for (int d = 0; d < 3; d++) I.prd[d] = E->prd[d];
  CK(cudaMemcpyAsync(E->imgp.p, &I, sizeof I, cudaMemcpyHostToDevice, st));
  // unsupported constellation anywhere: every rank keeps the pack-kernel path
  CK(cudaMemcpyAsync(hc, cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int ok = hc[2] ? 0 : 1;
  if (!solo) {
    E->hcnt[0] = ok;
    CK(cudaMemcpyAsync(E->cnt_dev.p, E->hcnt, sizeof(int), cudaMemcpyHostToDevice, st));
    NK(g_nccl.AllReduce(E->cnt_dev.p, E->cnt_dev.p + 1, 1, ncclInt, ncclMin, E->comm, st));
    CK(cudaMemcpyAsync(E->hcnt + 1, E->cnt_dev.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ok = E->hcnt[1];
  }
  if (!ok) return;
  E->fz_rq[0] = rq[0]; E->fz_rq[1] = rq[1];
  E->fz_ready = solo ? 2 : 1;
}

// flags -> deterministic compact list; returns the number of set flags
static int compact_flags(dem_engine *E, int n, DevBuf<int> &flag, DevBuf<int> &scan, DevBuf<int> &list)
{
  cudaStream_t st = E->stream;
  if (n <= 0) return 0;
  size_t tb = E->cubtmp.n;
  CK(cub::DeviceScan::ExclusiveSum(E->cubtmp.p, tb, flag.p, scan.p, n, st));
  CK(cudaMemcpyAsync(&E->hcnt[60], flag.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&E->hcnt[61], scan.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int cnt = E->hcnt[60] + E->hcnt[61];
  if (cnt) { list.ensure(E, cnt + 1); k_compact<<<GRID(n, 256), 256, 0, st>>>(n, flag.p, scan.p, list.p); }
  E->launches += 2;
  return cnt;
}

// per-step ghost refresh == CommBrick::forward_comm (comm_brick.cpp:563-645): one pack kernel per swap; periodic
// images on the same rank are written in place, remote ghosts travel as three NCCL send/recv pairs that land
// directly in the ghost region of the record arrays.  Swaps run in order so that edge/corner ghosts propagate.
static int *flag_slot(dem_engine *E, int slot);
static void do_swap(dem_engine *E, dem_engine::Swap *Ws, int nsw, bool allow_p2p = false, bool with_flags = false, int flags_slot = 0)
{  // nsw = 1, or the two (independent) swaps of one decomposed dimension exchanged in ONE NCCL group
  cudaStream_t st = E->stream;
  const int c = E->cur;
  if (allow_p2p && E->p2p_ok && !Ws[0].self) {  // peer-memory path (between rebuilds): one push launch, one wait launch
    const bool skip = E->opt.count("debug") && ((int)E->opt["debug"] & 8);
    PushP Q; memset(&Q, 0, sizeof Q);
    WaitP Wt; memset(&Wt, 0, sizeof Wt);
    Q.nsw = nsw;
    for (int q = 0; q < nsw; q++) {
      dem_engine::Swap &W = Ws[q];
      const int qi = (int)(&W - E->swaps);
      W.serial++;
      const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1), peer_recv = neighbor_rank(E, W.dim, W.side ? -1 : 1);
      if (peer_send >= 0 && !skip) {
        const int pc = W.pcur0 ^ (c ^ E->cur0);  // the receiver's buffer parity moves in lock step with mine
        SwapP &S = Q.S[q];
        S.n = W.nsend; S.list = W.list.p; S.dim = W.dim; S.shift = W.shift; S.xr = E->xr[c].p; S.vm = E->vm[c].p; S.wt = E->wt[c].p;
        S.ox = W.pbase[0 + pc] + W.pgfirst; S.ov = W.pbase[2 + pc] + W.pgfirst; S.ow = W.pbase[4 + pc] + W.pgfirst;
        Q.nb[q] = (int)GRID(W.nsend, 256); Q.done[q] = (unsigned *)E->hsig.p + 8 + qi; Q.sig[q] = W.psig + qi; Q.serial[q] = W.serial;
      }
      if (peer_recv >= 0 && !skip) { Wt.sig[q] = E->hsig.p + qi; Wt.serial[q] = W.serial; }
    }
    if (with_flags) {
      E->fserial++;
      Q.with_flags = 1; Q.myflags = flag_slot(E, flags_slot); Q.me = E->rank; Q.nranks = E->nranks; Q.slot = flags_slot; Q.fserial = E->fserial;
      for (int r = 0; r < E->nranks; r++) Q.peer_box[r] = E->peer_fbox[r];
      Wt.with_flags = 1; Wt.box = E->fbox.p; Wt.nranks = E->nranks; Wt.slot = flags_slot; Wt.fserial = E->fserial;
      Wt.gate_out = flag_slot(E, flags_slot) + 4; Wt.host_out = E->hflag_dev + 4 * flags_slot; Wt.host_serial = E->hflag_dev + 16 + flags_slot;
      E->fser_slot[flags_slot] = E->fserial;
    }
    Wt.timeout_flag = E->hsig.p + 15;
    if (with_flags) Wt.zero_next = flag_slot(E, flags_slot ^ 1);
    k_push<<<(unsigned)(Q.nb[0] + Q.nb[1] + 1), 256, 0, st>>>(Q);
    k_wait<<<1, 32, 0, st>>>(Wt);
    E->launches += 2;
    return;
  }
  size_t off[2] = {0, 0}, tot = 0;
  for (int q = 0; q < nsw; q++) { off[q] = tot; if (!Ws[q].self) tot += 3 * (size_t)Ws[q].nsend; }
  if (tot) E->sbuf.ensure(E, tot, 0, st);
  bool any_remote = false;
  for (int q = 0; q < nsw; q++) {
    dem_engine::Swap &W = Ws[q];
    SwapP S;
    S.n = W.nsend; S.list = W.list.p; S.dim = W.dim; S.shift = W.shift; S.xr = E->xr[c].p; S.vm = E->vm[c].p; S.wt = E->wt[c].p;
    if (W.self) { S.ox = E->xr[c].p + W.gfirst; S.ov = E->vm[c].p + W.gfirst; S.ow = E->wt[c].p + W.gfirst; }
    else { any_remote = true; S.ox = E->sbuf.p + off[q]; S.ov = S.ox + W.nsend; S.ow = S.ox + 2 * (size_t)W.nsend; }
    if (W.nsend) { k_pack_swap<<<GRID(W.nsend, 256), 256, 0, st>>>(S); E->launches++; }
  }
  if (!any_remote) return;
  if (E->opt.count("debug") && ((int)E->opt["debug"] & 8) && E->setup_done) return;  // profiling aid: no halo traffic between rebuilds
  NK(g_nccl.GroupStart());
  for (int q = 0; q < nsw; q++) {
    dem_engine::Swap &W = Ws[q];
    if (W.self) continue;
    const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1), peer_recv = neighbor_rank(E, W.dim, W.side ? -1 : 1);
    if (peer_send >= 0 && W.nsend)
      for (int a = 0; a < 3; a++) NK(g_nccl.Send(E->sbuf.p + off[q] + (size_t)a * W.nsend, 4 * (size_t)W.nsend, ncclDouble, peer_send, E->comm, st));
    if (peer_recv >= 0 && W.nrecv) {
      NK(g_nccl.Recv(E->xr[c].p + W.gfirst, 4 * (size_t)W.nrecv, ncclDouble, peer_recv, E->comm, st));
      NK(g_nccl.Recv(E->vm[c].p + W.gfirst, 4 * (size_t)W.nrecv, ncclDouble, peer_recv, E->comm, st));
      NK(g_nccl.Recv(E->wt[c].p + W.gfirst, 4 * (size_t)W.nrecv, ncclDouble, peer_recv, E->comm, st));
    }
  }
  NK(g_nccl.GroupEnd());
}
static bool peer_flags(const dem_engine *E) { return E->nranks > 1 && E->p2p_ok && E->fbox_ready == 1; }
// ghost refresh of a step; flags_slot >= 0: the step's flags (device slot flags_slot) travel with the last decomposed
// dimension's push when the peer-memory flag boxes are in use (returns true: post_flags has nothing left to do)
static bool forward_comm(dem_engine *E, int flags_slot = -1)
{
  // the two swaps of a dimension never feed each other (comm_brick.cpp:899-905: the candidates of both are the atoms
  // present before the dimension starts), so they travel together; dimensions stay ordered (edge / corner ghosts)
  int last_remote = -1;
  for (int q = 0; q < E->nswap; q++) if (!E->swaps[q].self) last_remote = q;
  const bool pf = flags_slot >= 0 && peer_flags(E) && last_remote >= 0;
  bool sent = false;
  for (int q = 0; q < E->nswap;) {
    const int n = (q + 1 < E->nswap && E->swaps[q + 1].dim == E->swaps[q].dim) ? 2 : 1;
    const bool wf = pf && last_remote >= q && last_remote < q + n;
    do_swap(E, &E->swaps[q], n, true, wf, flags_slot);
    sent = sent || wf;
    q += n;
  }
  return sent;
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

// ---- particle insertion between two runs (what `create_atoms`, `fix insert/pack|stream` do to the path's state:
// atom.cpp / atom_vec_sphere.cpp create_atom appends the new atoms behind the owned ones with zero force, no partners, no
// wall history; the next run's Verlet::setup rebuilds the lists and keeps the history of the existing contacts).  The new
// particles are appended behind the owned ones on the device; the old neighbour list gets empty rows for them so that the
// rebuild's history remap (by partner tag) carries every existing contact over.
static void grow_list_rows(dem_engine *E, ListSet &L, int ncap)
{  // re-stride the [row][cap] arrays of a list to a larger capacity, contents kept
  if (ncap <= L.cap) return;
  cudaStream_t st = E->stream;
  DevBuf<unsigned> nbr; DevBuf<int> ptag, nn; DevBuf<double4> hist;
  nbr.ensure(E, (size_t)L.maxk * ncap); ptag.ensure(E, (size_t)L.maxk * ncap); nn.ensure(E, ncap);
  CK(cudaMemsetAsync(nn.p, 0, (size_t)ncap * sizeof(int), st));
  CK(cudaMemcpy2DAsync(nbr.p, (size_t)ncap * sizeof(unsigned), L.nbr.p, (size_t)L.cap * sizeof(unsigned), (size_t)L.cap * sizeof(unsigned), L.maxk, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpy2DAsync(ptag.p, (size_t)ncap * sizeof(int), L.ptag.p, (size_t)L.cap * sizeof(int), (size_t)L.cap * sizeof(int), L.maxk, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(nn.p, L.numneigh.p, (size_t)L.cap * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (L.dnum) {
    hist.ensure(E, (size_t)L.hslots * L.dnum * ncap);
    CK(cudaMemcpy2DAsync(hist.p, (size_t)ncap * sizeof(double4), L.hist.p, (size_t)L.cap * sizeof(double4), (size_t)L.cap * sizeof(double4), (size_t)L.hslots * L.dnum, cudaMemcpyDeviceToDevice, st));
  }
  CK(cudaStreamSynchronize(st));
  L.nbr.release(); L.ptag.release(); L.numneigh.release(); L.hist.release();
  L.nbr = nbr; L.ptag = ptag; L.numneigh = nn; L.hist = hist; L.cap = ncap;
}

extern "C" int dem_insert_particles(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x,
                                    const double *v, const double *omega, const double *radius, const double *density)
{
  if (e && !e->uploaded) return dem_upload_particles(e, n, tag, type, mask, x, v, omega, radius, density);  // the first particles of the deck
  API_BEGIN
  if (n < 0 || (n > 0 && (!tag || !type || !x || !radius || !density))) dem_fail(e, DEM_ERR_ARG, "missing particle arrays");
  if (n == 0) return DEM_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  // this rank's share (several bricks: ownership test of dem_upload_particles)
  std::vector<long> mine; mine.reserve(n);
  const MineP B = brick_params(e);
  double rmax_all = 0.0, rmin_all = 1e300;
  for (long i = 0; i < n; i++) {
    if (type[i] < 1 || type[i] > e->ntypes) dem_fail(e, DEM_ERR_ARG, "Invalid atom type in particle data");
    if (!(radius[i] > 0.0) || !(density[i] > 0.0)) dem_fail(e, DEM_ERR_ARG, "Invalid radius or density in particle data");
    if (tag[i] <= 0) dem_fail(e, DEM_ERR_ARG, "Invalid atom ID in particle data");
    rmax_all = std::max(rmax_all, radius[i]); rmin_all = std::min(rmin_all, radius[i]);
    if (e->nranks == 1 || brick_owns(B, x + 3 * i)) mine.push_back(i);
  }
  const long nm = (long)mine.size(), n0 = e->nlocal, need = n0 + nm;
  if (need >= (long)NBR_IDX) dem_fail(e, DEM_ERR_OVERFLOW, "more than 2^25 particles on one GPU");
  if (rmax_all > e->rmax || rmin_all < e->rmin || !(e->rmin > 0.0)) {  // the neighbour cutoff / cell grid follow the extreme radii
    e->rmax = std::max(e->rmax, rmax_all); e->rmin = e->rmin > 0.0 ? std::min(e->rmin, rmin_all) : rmin_all; e->dirty = 1;
  }
  if (nm) {
    if (need > e->cap) ensure_particle_cap(e, need + need / 4 + 1024, n0);
    std::vector<double> hx(3 * nm), hv(3 * nm, 0.0), hw(3 * nm, 0.0), hr(nm), hd(nm);
    std::vector<int> ht(nm), hm(nm, 1), hg(nm);
    for (long q = 0; q < nm; q++) {
      const long i = mine[q];
      for (int d = 0; d < 3; d++) { hx[3 * q + d] = x[3 * i + d]; if (v) hv[3 * q + d] = v[3 * i + d]; if (omega) hw[3 * q + d] = omega[3 * i + d]; }
      hr[q] = radius[i]; hd[q] = density[i]; ht[q] = type[i]; if (mask) hm[q] = mask[i]; hg[q] = tag[i];
    }
    const size_t nd = (size_t)nm;
    e->stage.ensure(e, nd * 11 * sizeof(double) + nd * 3 * sizeof(int) + 64);
    double *dx = (double *)e->stage.p, *dv = dx + 3 * nd, *dw = dv + 3 * nd, *dr = dw + 3 * nd, *dd = dr + nd;
    int *dt = (int *)(dd + nd), *dm = dt + nd, *dg = dm + nd;
    CK(cudaMemcpyAsync(dx, hx.data(), 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dv, hv.data(), 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dw, hw.data(), 3 * nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dr, hr.data(), nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dd, hd.data(), nd * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dt, ht.data(), nd * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dm, hm.data(), nd * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dg, hg.data(), nd * sizeof(int), cudaMemcpyHostToDevice, st));
    e->counters.ensure(e, 4);
    CK(cudaMemsetAsync(e->counters.p, 0, 2 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(e->counters.p + 2, 0xFF, sizeof(unsigned long long), st));
    const int c = e->cur;
    k_pack_upload<<<GRID(nm, 256), 256, 0, st>>>((int)nm, dx, dv, dw, dr, dd, dt, dm, dg, e->ntypes, e->xr[c].p + n0, e->vm[c].p + n0, e->wt[c].p + n0,
                                                  (int *)(e->counters.p + 1), e->counters.p, e->ins_mass);
    e->launches++;
    CK(cudaMemcpyAsync(e->tag.p + n0, dg, nd * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(e->density.p + n0, dd, nd * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(e->xh.p + n0, 0, nd * sizeof(double4), st));  // no wall candidate / history-valid bits
    if (e->f.p) for (int d = 0; d < 3; d++) { CK(cudaMemsetAsync(e->f.p + (size_t)d * e->cap + n0, 0, nd * sizeof(double), st)); CK(cudaMemsetAsync(e->tq.p + (size_t)d * e->cap + n0, 0, nd * sizeof(double), st)); }
    if (e->whist.p) for (int r = 0; r < e->nwrows; r++) CK(cudaMemsetAsync(e->whist.p + (size_t)r * e->cap + n0, 0, nd * sizeof(double), st));
    if (e->mesh_ready) for (int b = 0; b < 2; b++) if (e->mint[b].p) CK(cudaMemsetAsync(e->mint[b].p + n0, 0, nd * sizeof(int), st));  // no mesh contact rows
    for (int b = 0; b < 2; b++) {  // the old list: empty rows for the newcomers
      ListSet &L = e->ls[b];
      if (!L.valid) continue;
      if (need > L.cap) { if (L.dnum) grow_list_rows(e, L, e->cap); else { L.valid = 0; continue; } }
      CK(cudaMemsetAsync(L.numneigh.p + n0, 0, nd * sizeof(int), st));
    }
    CK(cudaStreamSynchronize(st));
    e->nlocal = need;
  }
  e->nghost = 0;
  e->forces_valid = 0; e->order_valid = 0; e->need_setup = 1;
  API_END
}

// particle migration between bricks: CommBrick::exchange (comm_brick.cpp:732-860); runs BEFORE the periodic
// wrap so that the direction is decided on the unwrapped coordinate (the sender wraps while packing)
static int migrate(dem_engine *E, int ncur, int &ngone)
{
  cudaStream_t st = E->stream;
  const int c = E->cur;
  ListSet &Lold = E->ls[E->lcur];
  const bool hist = Lold.valid && Lold.dnum > 0;
  const int hrec = hist ? Lold.dnum : 0;
  ngone = 0;
  E->gone.ensure(E, E->cap); CK(cudaMemsetAsync(E->gone.p, 0, (size_t)E->cap * sizeof(int), st));
  for (int d = 0; d < 3; d++) {
    if (E->pgrid[d] == 1) continue;
    const int plo = neighbor_rank(E, d, -1), phi = neighbor_rank(E, d, 1);
    E->flo.ensure(E, ncur + 1); E->fhi.ensure(E, ncur + 1); E->slo.ensure(E, ncur + 1); E->shi.ensure(E, ncur + 1);
    if (ncur) k_mig_flag<<<GRID(ncur, 256), 256, 0, st>>>(ncur, E->xr[c].p, d, E->sublo[d], E->subhi[d], plo >= 0, phi >= 0, E->flo.p, E->fhi.p, E->gone.p);
    E->launches++;
    DevBuf<int> lists[2];
    int nsend[2] = {compact_flags(E, ncur, E->flo, E->slo, lists[0]), 0};
    nsend[1] = compact_flags(E, ncur, E->fhi, E->shi, lists[1]);
    for (int side = 0; side < 2; side++) {
      const int peer_send = side ? phi : plo, peer_recv = side ? plo : phi;
      int H = 0;
      if (hist && nsend[side]) {
        CK(cudaMemsetAsync(E->overflow.p, 0, sizeof(int), st));
        if (Lold.fmt) k_max_nh2<<<GRID(nsend[side], 256), 256, 0, st>>>(nsend[side], lists[side].p, Lold.numneigh.p, Lold.nbr.p, Lold.cap, Lold.maxk, E->overflow.p);
        else k_max_nh<<<GRID(nsend[side], 256), 256, 0, st>>>(nsend[side], lists[side].p, Lold.numneigh.p, E->overflow.p);
        CK(cudaMemcpyAsync(&E->hcnt[62], E->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        H = E->hcnt[62];
      }
      int sv[2] = {nsend[side], H}, rv[2] = {0, 0};
      xchg_ints(E, peer_send, sv, peer_recv, rv, 2);
      const int nrecv = rv[0], Hr = rv[1];
      MigP M;
      memset(&M, 0, sizeof M);
      M.nwrows = E->nwrows; M.hrec = hrec; M.cap = E->cap; M.lcap = Lold.cap; M.dim = d;
      M.periodic = E->periodic[d]; M.wrap_lo = E->lo[d]; M.wrap_hi = E->hi[d]; M.prd = E->prd[d];
      M.density = E->density.p; M.whist = E->whist.p;
      if (hist) { M.nbr = Lold.nbr.p; M.numneigh = Lold.numneigh.p; M.ptag = Lold.ptag.p; M.hist = Lold.hist.p; M.hslots = Lold.hslots; M.maxk = Lold.maxk; }
      const bool meshrows = E->mesh_ready && have_mesh_walls(E);
      const int mblock = meshrows ? E->mslots * (1 + 4 * E->mhrec) : 0;
      if (meshrows) { M.mslots = E->mslots; M.mhrec = E->mhrec; }
      const int ss = 16 + E->nwrows + mblock + H * (1 + 4 * hrec), sr = 16 + E->nwrows + mblock + Hr * (1 + 4 * hrec);
      if (nsend[side]) {
        E->migs.ensure(E, (size_t)nsend[side] * ss);
        M.n = nsend[side]; M.stride = ss; M.hmax = H; M.list = lists[side].p; M.buf = E->migs.p;
        M.xr = E->xr[c].p; M.vm = E->vm[c].p; M.wt = E->wt[c].p; M.xh = E->xh.p; M.tag = E->tag.p;
        if (meshrows) { M.mint = E->mint[E->mcur].p; M.mhist = E->mhist[E->mcur].p; }
        if (hist && Lold.fmt) k_mig_pack2<<<GRID(M.n, 128), 128, 0, st>>>(M, Lold.nlocal);
        else k_mig_pack<<<GRID(M.n, 128), 128, 0, st>>>(M);
        E->launches++;
      }
      if (nrecv) {
        if ((long)ncur + nrecv > E->cap) ensure_particle_cap(E, (long)ncur + nrecv, ncur);
        if (hist && (ncur + nrecv > Lold.cap || Hr > Lold.maxk || Hr > Lold.hslots))
          dem_fail(E, DEM_ERR_OVERFLOW, "migration does not fit the neighbour rows (raise options cap_factor / maxneigh / histslots)");
        E->migr.ensure(E, (size_t)nrecv * sr);
      }
      NK(g_nccl.GroupStart());
      if (peer_send >= 0 && nsend[side]) NK(g_nccl.Send(E->migs.p, (size_t)nsend[side] * ss, ncclDouble, peer_send, E->comm, st));
      if (peer_recv >= 0 && nrecv) NK(g_nccl.Recv(E->migr.p, (size_t)nrecv * sr, ncclDouble, peer_recv, E->comm, st));
      NK(g_nccl.GroupEnd());
      if (nrecv) {
        E->gone.ensure(E, E->cap, ncur, st);
        CK(cudaMemsetAsync(E->gone.p + ncur, 0, (size_t)nrecv * sizeof(int), st));
        M.n = nrecv; M.stride = sr; M.hmax = Hr; M.buf = E->migr.p; M.cap = E->cap;
        M.xr = E->xr[E->cur].p; M.vm = E->vm[E->cur].p; M.wt = E->wt[E->cur].p; M.xh = E->xh.p; M.tag = E->tag.p;
        M.density = E->density.p; M.whist = E->whist.p;
        if (meshrows) { M.mint = E->mint[E->mcur].p; M.mhist = E->mhist[E->mcur].p; }  // (re-strided if the capacity grew)
        if (hist && Lold.fmt) k_mig_unpack2<<<GRID(nrecv, 128), 128, 0, st>>>(M, ncur);
        else k_mig_unpack<<<GRID(nrecv, 128), 128, 0, st>>>(M, ncur);
        E->launches++;
        ncur += nrecv;
      }
      ngone += nsend[side];
    }
    lists[0].release(); lists[1].release();
  }
  return ncur;
}

// Neighbor rebuild: verlet.cpp:305-328 (pre_exchange .. neighbor->build) re-designed for the GPU.
static void rebuild(dem_engine *E)
{
  cudaStream_t st = E->stream;
  StageTrace tr(st);
  const int dnum = E->have_pair ? E->pm.hrec : 0;  // history records per contact
  ensure_cub(E, (size_t)E->cap);
  E->overflow.ensure(E, 2);
  BoxP B;
  for (int d = 0; d < 3; d++) { B.lo[d] = E->lo[d]; B.hi[d] = E->hi[d]; B.prd[d] = E->prd[d]; B.periodic[d] = E->periodic[d]; }
  // 0. migration between bricks (multi-rank only)
  int ncur = (int)E->nlocal, ngone = 0;
  if (E->nranks > 1) ncur = migrate(E, ncur, ngone);
  const int n = ncur - ngone;
  ensure_cub(E, (size_t)E->cap);
  E->keys.ensure(E, E->cap); E->keys2.ensure(E, E->cap); E->vals.ensure(E, E->cap); E->perm.ensure(E, E->cap);
  int c = E->cur;
  bool did_permute = false;
  if (ncur) {
    // 1. pbc wrap, cell keys, radix sort, gather the particle records into cell (Morton) order
    k_wrap_key<<<GRID(ncur, 256), 256, 0, st>>>(ncur, E->xr[c].p, E->grid, B, E->keys.p, E->vals.p, E->nranks > 1 ? E->gone.p : nullptr);
    size_t tb = E->cubtmp.n;
    CK(cub::DeviceRadixSort::SortPairs(E->cubtmp.p, tb, E->keys.p, E->keys2.p, E->vals.p, E->perm.p, ncur, 0, 32, st));
    E->launches += 2;
  }
  if (n) {
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
    E->launches += 4;
    E->cur = c ^ 1; c = E->cur;
    did_permute = true;
  }
  tr.mark("migrate+sort+gather");
  E->nlocal = n;
  // 2. owned cell ranges
  CK(cudaMemsetAsync(E->ocs.p, 0, E->ncells * sizeof(int), st)); CK(cudaMemsetAsync(E->oce.p, 0, E->ncells * sizeof(int), st));
  CK(cudaMemsetAsync(E->gcs.p, 0, E->ncells * sizeof(int), st)); CK(cudaMemsetAsync(E->gce.p, 0, E->ncells * sizeof(int), st));
  if (n) { k_cell_ranges<<<GRID(n, 256), 256, 0, st>>>(n, 0, E->xr[c].p, E->grid, E->ocs.p, E->oce.p); E->launches++; }
  // 3. ghosts: CommBrick::borders (comm_brick.cpp:884-1117).  One dimension after the other, two swaps per
  // dimension, candidates = owned + ghosts of earlier dimensions, so edges and corners propagate.
  E->nghost = 0; E->nswap = 0;
  for (int d = 0; d < 3; d++) {
    const bool multi = E->pgrid[d] > 1;
    if (!multi && !E->periodic[d]) continue;
    if (!multi && n == 0) continue;
    const int plo = multi ? neighbor_rank(E, d, -1) : E->rank, phi = multi ? neighbor_rank(E, d, 1) : E->rank;
    const int n0 = (int)(E->nlocal + E->nghost);
    E->flo.ensure(E, n0 + 1); E->fhi.ensure(E, n0 + 1); E->slo.ensure(E, n0 + 1); E->shi.ensure(E, n0 + 1);
    // no neighbour on a side: the flag can never be set (cut beyond the data)
    const double lo_cut = plo >= 0 ? E->sublo[d] + E->cutneighmax : -1e300, hi_cut = phi >= 0 ? E->subhi[d] - E->cutneighmax : 1e300;
    if (n0) k_border_flag<<<GRID(n0, 256), 256, 0, st>>>(n0, E->xr[c].p, d, lo_cut, hi_cut, E->flo.p, E->fhi.p);
    E->launches++;
    for (int side = 0; side < 2; side++) {
      dem_engine::Swap &W = E->swaps[E->nswap];
      W.dim = d; W.side = side; W.self = multi ? 0 : 1;
      W.nsend = compact_flags(E, n0, side ? E->fhi : E->flo, side ? E->shi : E->slo, W.list);
      W.shift = 0.0;
      if (E->periodic[d] && side == 0 && E->myloc[d] == 0) W.shift = E->prd[d];
      if (E->periodic[d] && side == 1 && E->myloc[d] == E->pgrid[d] - 1) W.shift = -E->prd[d];
      const int peer_send = side ? phi : plo, peer_recv = side ? plo : phi;
      if (multi) { int sv[1] = {W.nsend}, rv[1] = {0}; xchg_ints(E, peer_send, sv, peer_recv, rv, 1); W.nrecv = rv[0]; }
      else W.nrecv = W.nsend;
      W.gfirst = (int)(E->nlocal + E->nghost);
      if (W.gfirst + W.nrecv > E->cap) ensure_particle_cap(E, (long)W.gfirst + W.nrecv, W.gfirst);
      c = E->cur;
      // tags of the new ghosts
      if (W.self) { if (W.nsend) k_pack_int<<<GRID(W.nsend, 256), 256, 0, st>>>(W.nsend, W.list.p, E->tag.p, E->tag.p + W.gfirst); }
      else {
        if (W.nsend) { E->sbuf_i.ensure(E, W.nsend); k_pack_int<<<GRID(W.nsend, 256), 256, 0, st>>>(W.nsend, W.list.p, E->tag.p, E->sbuf_i.p); }
        NK(g_nccl.GroupStart());
        if (peer_send >= 0 && W.nsend) NK(g_nccl.Send(E->sbuf_i.p, W.nsend, ncclInt, peer_send, E->comm, st));
        if (peer_recv >= 0 && W.nrecv) NK(g_nccl.Recv(E->tag.p + W.gfirst, W.nrecv, ncclInt, peer_recv, E->comm, st));
        NK(g_nccl.GroupEnd());
      }
      E->launches++;
      E->nghost += W.nrecv;
      E->nswap++;
    }
    // records of this dimension's new ghosts (the next dimension's flags look at them)
    do_swap(E, &E->swaps[E->nswap - 2], 2);
  }
  if (E->nlocal + E->nghost >= (long)NBR_IDX)
    dem_fail(E, DEM_ERR_OVERFLOW, "%ld owned + %ld ghost particles exceed the 2^25 indices of a neighbour word", E->nlocal, E->nghost);
  tr.mark("cell ranges + borders");
  // 4. cell order of the ghosts (storage keeps the swap order so that NCCL can receive in place)
  if (E->nghost) {
    const int ng = (int)E->nghost;
    E->keys.ensure(E, E->cap); E->keys2.ensure(E, E->cap); E->vals.ensure(E, E->cap); E->gorder.ensure(E, E->cap);
    ensure_cub(E, (size_t)E->cap);
    k_ghost_keys<<<GRID(ng, 256), 256, 0, st>>>(ng, n, E->xr[c].p, E->grid, E->keys.p, E->vals.p);
    size_t tb = E->cubtmp.n;
    CK(cub::DeviceRadixSort::SortPairs(E->cubtmp.p, tb, E->keys.p, E->keys2.p, E->vals.p, E->gorder.p, ng, 0, 32, st));
    k_cell_ranges_idx<<<GRID(ng, 256), 256, 0, st>>>(ng, E->gorder.p, E->xr[c].p, E->grid, E->gcs.p, E->gce.p);
    E->launches += 3;
  }
  tr.mark("ghost order");
  // 5. full Verlet list + history remap
  ListSet &Lold = E->ls[E->lcur], &Lnew = E->ls[E->lcur ^ 1];
  // full list (k_step, k_step_bond) unless option owner_list asks for the measured alternative of dem_pairs.cuh
  const int fmt = (E->have_pair && !E->pm.cohesion && E->pm.normal < N_HYST1 && E->cdf == 1.0 && E->opt.count("owner_list") && E->opt["owner_list"] != 0) ? 1 : 0;
  int maxk = std::max(Lold.valid ? Lold.maxk : 0, (int)(E->opt.count("maxneigh") ? E->opt["maxneigh"] : 24));
  // history rows: by default one per row entry (a particle can never gain more contacts between two rebuilds than it has
  // list entries, so the step kernels cannot run out of rows and drop a contact's history; rows that are not in use cost
  // address space, not traffic).  Option histslots lowers it for memory-tight runs; an overflow is then reported by dem_run.
  int hslots = std::max(Lold.valid ? Lold.hslots : 0, (int)(E->opt.count("histslots") ? E->opt["histslots"] : std::min(maxk, NBR_MAXSLOTS)));
  for (int attempt = 0; attempt < 6; attempt++) {
    ensure_list(E, Lnew, E->cap, maxk, dnum, hslots);
    Lnew.fmt = fmt; Lnew.nlocal = n;
    if (!n) break;
    CK(cudaMemsetAsync(E->overflow.p, 0, 2 * sizeof(int), st));
    BuildP P;
    P.nlocal = n; P.cap = Lnew.cap; P.maxk = Lnew.maxk; P.dnum = dnum; P.hslots = Lnew.hslots; P.xr = E->xr[c].p; P.tag = E->tag.p; P.G = E->grid;
    P.ocs = E->ocs.p; P.oce = E->oce.p; P.gcs = E->gcs.p; P.gce = E->gce.p; P.gorder = E->gorder.p; P.cdf = E->cdf; P.skin = E->skin;
    P.nbr = Lnew.nbr.p; P.numneigh = Lnew.numneigh.p; P.ptag = Lnew.ptag.p; P.hist = Lnew.hist.p;
    P.have_old = (Lold.valid && dnum && Lold.dnum == dnum && Lold.fmt == fmt) ? 1 : 0; P.cap_old = Lold.cap; P.dnum_old = Lold.dnum;
    P.perm = E->perm.p; P.nbr_old = Lold.nbr.p; P.numneigh_old = Lold.numneigh.p; P.ptag_old = Lold.ptag.p; P.hist_old = Lold.hist.p;
    P.overflow = E->overflow.p;
    P.coh_nbond = (E->have_pair && E->pm.cohesion) ? E->pm.nbond : 0; P.coh_rec = E->pm.rec_bond;
    if (fmt) {
      k_build_list2<<<GRID(n, 128), 128, 0, st>>>(P, Lold.nlocal, Lold.maxk);
      k_link_slots<<<GRID(n, 128), 128, 0, st>>>(n, Lnew.cap, Lnew.maxk, Lnew.nbr.p, Lnew.numneigh.p);
      E->launches++;
    } else k_build_list<<<GRID(n, 128), 128, 0, st>>>(P);
    E->launches++;
    int ov[2] = {0, 0};
    CK(cudaMemcpyAsync(ov, E->overflow.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ov[0] == 0 && ov[1] == 0) break;
    if (attempt == 5) dem_fail(E, DEM_ERR_OVERFLOW, "neighbour list overflow (%d neighbours, %d history partners)", ov[0], ov[1]);
    if (ov[0]) { maxk = ov[0] + 4; if (!E->opt.count("histslots")) hslots = std::max(hslots, std::min(maxk, NBR_MAXSLOTS)); }
    if (ov[1]) hslots = std::max(hslots, ov[1] + 12);
    if (maxk > (fmt ? NN2_MAXK : 0xffff) || hslots > NBR_MAXSLOTS) dem_fail(E, DEM_ERR_OVERFLOW, "a particle has %d neighbours / %d history partners", ov[0], ov[1]);
  }
  if (fmt && E->res.n < (size_t)Lnew.hslots * 2 * Lnew.cap) { E->res.release(); E->res.ensure(E, (size_t)Lnew.hslots * 2 * Lnew.cap); }
  if (!fmt && E->opt.count("contact_output") && E->opt["contact_output"] != 0 && E->cout.n < (size_t)Lnew.hslots * 2 * Lnew.cap) {
    E->cout.release(); E->cout.ensure(E, (size_t)Lnew.hslots * 2 * Lnew.cap);
  }
  Lnew.valid = 1; Lold.valid = 0;
  E->lcur ^= 1;
  tr.mark("list build + remap");
  // 5b. triangle-mesh candidate rows + carry-over of the mesh contact rows
  const bool meshw = have_mesh_walls(E) && E->mesh_ready;
  if (meshw) mesh_rebuild(E, n, did_permute);
  // 6. positions at build time + primitive wall candidate bits
  E->nwc = 0;
  if (n) {
    const int nw = (int)E->walls.size();
    E->flo.ensure(E, n + 1); E->slo.ensure(E, n + 1);
    k_hold<<<GRID(n, 256), 256, 0, st>>>(n, E->xr[c].p, E->xh.p, E->valid_tmp.p, E->dwalls.p, nw, E->skin, (nw || meshw) ? E->flo.p : nullptr,
                                       meshw ? E->mint[E->mcur].p : nullptr);
    E->launches++;
    if (nw || meshw) {
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
  tr.mark("mesh + hold + wall index");
  halo_p2p_setup(E);
  tr.mark("halo p2p setup");
  fused_halo_setup(E);
  tr.mark("fused halo setup");
  E->ago = 0;
  E->order_valid = 0;
  E->nbuilds++;
}

static int *flag_slot(dem_engine *E, int slot);
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
  P.res = E->res.p; P.serial = (double)E->serial;
  P.whist = E->whist.p; P.f = E->f.p; P.tq = E->tq.p; P.walls = E->dwalls.p; P.nwalls = (int)E->walls.size(); P.nwc = E->nwc; P.nwcap = E->nwcap; P.wlist = E->wlist.p; P.fw = E->fw.p;
  P.pm = E->pm; P.tab = E->tab.p; P.nt1 = E->ntypes + 1;
  for (int w = 0; w < T_B_LAMBDA; w++) P.t1[w] = E->t1[w];
  P.ntimestep = E->ntimestep; P.tsCreateBond = (long)(int)E->tsCreateBond;
  P.dt = E->dt; P.dtv = E->dt; P.dtf = 0.5 * E->dt * E->ftm2v; P.dtfrot = P.dtf / 0.4;  // fix_nve.cpp:86, fix_nve_sphere.cpp:69,150
  P.nktv2p = E->nktv2p; P.charVel = E->charVel; P.cdf = E->cdf; P.cdfsq = E->cdf * E->cdf;
  P.trigsq = 0.25 * E->skin * E->skin;  // neighbor.cpp:298
  P.cutneighmax = E->cutneighmax;
  for (int d = 0; d < 3; d++) P.g[d] = E->g[d];
  P.have_g = E->have_g; P.have_pair = E->have_pair; P.freezebit = E->freezebit; P.integbit = E->integbit;
  P.nxf = (int)E->xf.size(); for (int q = 0; q < P.nxf; q++) P.xf[q] = E->xf[q];
  P.mode = mode; P.debug = E->opt.count("debug") ? (int)E->opt["debug"] : 0; P.flag = flag_slot(E, E->fslot); P.gate = E->gate; P.gate_mask = E->gate_mask; P.ncontact = nullptr;
  P.bondc = (E->have_pair && E->pm.cohesion == C_BOND) ? E->bondc.p : nullptr;
  P.img = nullptr; P.img_first = nullptr; P.img_tab = nullptr;
  if (E->fz_on) {  // fused ghost push: this launch writes buffer cur ^ 1
    P.img = E->imgp.p; P.img_first = E->img_first.p; P.img_tab = E->img_tab.p;
    P.img_par = (E->cur ^ 1) ^ E->cur0;
  }
  return P;
}

// result records are stamped with a serial number that is unique in the process (and starts at the wall clock in
// microseconds): device blocks are recycled between engines, so a stale record must never carry the current stamp
static long next_serial()
{
  static std::mutex mu; static long cur = 0;
  std::lock_guard<std::mutex> lk(mu);
  if (!cur) cur = (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count() & ((1L << 52) - 1);
  return ++cur;
}
template <int N, int R>
static void launch_pairs_t(dem_engine *E, const StepP &P)
{  // owner list: phase A (owned pairs, once each), then phase B (partner shares + integration)
  if (E->ntypes == 1) k_pairs<N, R, true><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
  else k_pairs<N, R, false><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
  k_finish<<<GRID(P.nlocal, 256), 256, 0, E->stream>>>(P);
  E->launches++;
}
template <int N, int R>
static void launch_step_t(dem_engine *E, const StepP &P)
{
  if (E->opt.count("fp32") && E->opt["fp32"] != 0) {  // fp32 mode: contact law in single precision (dem_contact.cuh pair_chain_f32)
    if (E->ntypes == 1) k_step<N, R, true, true><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
    else k_step<N, R, false, true><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
    return;
  }
  // the reference's default sub-model settings get the specialised instantiation (see pair_item)
  const ModelP &m = E->pm;
  const bool std_deck = E->have_pair && m.tangential && m.tdamp && !m.limitForce && !m.torsion && !m.cdtnl2 && P.nktv2p == 1.0 && P.cdf == 1.0 && !P.cout && !P.debug && !P.nxf &&
                        !(E->opt.count("generic_step") && E->opt["generic_step"] != 0);
  if (std_deck) {
    if (E->ntypes == 1) k_step<N, R, true, false, true><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
    else k_step<N, R, false, false, true><<<GRID(P.nlocal, 128), 128, 0, E->stream>>>(P);
    return;
  }
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
  if (mode != MODE_STEP && E->cout.p && E->ls[E->lcur].fmt == 0 && E->have_pair && !E->pm.cohesion) {
    E->cout_serial = next_serial(); P.cout = E->cout.p; P.serial = (double)E->cout_serial;
  }
  if (P.nwc) { k_walls<<<GRID(P.nwc, 128), 128, 0, E->stream>>>(P); E->launches++; }
  if (have_mesh_walls(E) && E->mesh_ready && E->any_stress) CK(cudaMemsetAsync(E->dmforce.p, 0, 6 * DEM_MAXMESH * sizeof(double), E->stream));  // MeshModuleStress::pre_force
  if (P.nwc && have_mesh_walls(E)) { MeshP M = mesh_params(E); mesh_launch_step(P, M, E->stream); E->launches++; }
  const int key = (E->have_pair ? E->pm.normal : N_HERTZ) * 4 + (E->have_pair ? E->pm.rolling : R_OFF);
  if (E->ls[E->lcur].fmt == 1) {
    if (P.nxf) dem_fail(E, DEM_ERR_UNSUPPORTED, "fix addforce / viscous with option owner_list");
    E->serial = next_serial(); P.serial = (double)E->serial;
    {
      switch (key) {
        case N_HERTZ * 4 + R_OFF: launch_pairs_t<N_HERTZ, R_OFF>(E, P); break;
        case N_HERTZ * 4 + R_CDT: launch_pairs_t<N_HERTZ, R_CDT>(E, P); break;
        case N_HERTZ * 4 + R_EPSD: launch_pairs_t<N_HERTZ, R_EPSD>(E, P); break;
        case N_HERTZ * 4 + R_EPSD2: launch_pairs_t<N_HERTZ, R_EPSD2>(E, P); break;
        case N_HOOKE * 4 + R_OFF: launch_pairs_t<N_HOOKE, R_OFF>(E, P); break;
        case N_HOOKE * 4 + R_CDT: launch_pairs_t<N_HOOKE, R_CDT>(E, P); break;
        case N_HOOKE * 4 + R_EPSD: launch_pairs_t<N_HOOKE, R_EPSD>(E, P); break;
        case N_HOOKE * 4 + R_EPSD2: launch_pairs_t<N_HOOKE, R_EPSD2>(E, P); break;
        default: dem_fail(E, DEM_ERR_STATE, "no kernel for this model combination");
      }
    }
  } else
  if (E->have_pair && E->pm.normal >= N_HYST1) {  // INL normal laws: general kernel (canonical orientation, 12 more history values per pair)
    const unsigned g = GRID(P.nlocal, 128);
#define HYSTK(R) { if (E->pm.normal == N_HYST1) k_step_hyst<N_HYST1, R><<<g, 128, 0, E->stream>>>(P); else k_step_hyst<N_HYST2, R><<<g, 128, 0, E->stream>>>(P); }
    switch (E->pm.rolling) { case R_OFF: HYSTK(R_OFF) break; case R_CDT: HYSTK(R_CDT) break; case R_EPSD: HYSTK(R_EPSD) break; default: HYSTK(R_EPSD2) break; }
#undef HYSTK
  } else
  if (E->have_pair && E->pm.cohesion) {
    const unsigned g = GRID(P.nlocal, 128);
#define BONDK(N, R) { if (E->pm.cohesion == C_BOND) k_step_bond<N, R, C_BOND><<<g, 128, 0, E->stream>>>(P); else k_step_bond<N, R, C_BONDNL><<<g, 128, 0, E->stream>>>(P); }
    switch (key) {
      case N_HERTZ * 4 + R_OFF: BONDK(N_HERTZ, R_OFF) break;
      case N_HERTZ * 4 + R_CDT: BONDK(N_HERTZ, R_CDT) break;
      case N_HERTZ * 4 + R_EPSD: BONDK(N_HERTZ, R_EPSD) break;
      case N_HERTZ * 4 + R_EPSD2: BONDK(N_HERTZ, R_EPSD2) break;
      case N_HOOKE * 4 + R_OFF: BONDK(N_HOOKE, R_OFF) break;
      case N_HOOKE * 4 + R_CDT: BONDK(N_HOOKE, R_CDT) break;
      case N_HOOKE * 4 + R_EPSD: BONDK(N_HOOKE, R_EPSD) break;
      default: BONDK(N_HOOKE, R_EPSD2) break;
    }
#undef BONDK
  } else
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

// Step flags ([0] rebuild trigger, [1] history overflow, [2] moving-mesh trigger) live in two device slots that alternate
// between steps: step n writes slot n&1 and is GATED by the slot step n-1 wrote (all-reduced over the ranks; the reference
// does MPI_Allreduce in neighbor.cpp:1463).  The host copy of a slot arrives one step late through an event, so the host
// queues step n before it knows whether step n must be preceded by a rebuild; if it must, the gated kernels of step n have
// returned without touching anything and the host redoes the step after the rebuild (see dem_run).
static int *flag_slot(dem_engine *E, int slot) { return E->dflag.p + 8 * slot; }
static const int *gate_slot(dem_engine *E, int slot) { return E->nranks > 1 ? E->dflag.p + 8 * slot + 4 : E->dflag.p + 8 * slot; }
static void clear_flags(dem_engine *E)
{
  E->dflag.ensure(E, 24);
  CK(cudaStreamSynchronize(E->stream));
  for (int k = 0; k < 8; k++) E->hflag[k] = 0;
  CK(cudaMemsetAsync(E->dflag.p, 0, 24 * sizeof(int), E->stream));
  E->slot_zeroed = 1;
}
// after a step: reduce its flags over the ranks, start the copy to the host, mark the point with an event
static void post_flags(dem_engine *E, int slot, bool sent_with_halo = false)
{
  if (sent_with_halo) { E->fev_peer[slot] = 1; return; }  // k_push / k_wait carried the flags: gate words and host block are written by k_wait
  E->fev_peer[slot] = 0;
  const int dbg = E->opt.count("debug") ? (int)E->opt["debug"] : 0;
  if (E->nranks > 1) {
    if (dbg & 4) CK(cudaMemcpyAsync(flag_slot(E, slot) + 4, flag_slot(E, slot), 4 * sizeof(int), cudaMemcpyDeviceToDevice, E->stream));  // profiling aid: no all-reduce
    else NK(g_nccl.AllReduce(flag_slot(E, slot), flag_slot(E, slot) + 4, 4, ncclInt, ncclMax, E->comm, E->stream));
  }
  CK(cudaMemcpyAsync(E->hflag + 4 * slot, gate_slot(E, slot), 4 * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
  if (!E->fev[slot]) CK(cudaEventCreateWithFlags(&E->fev[slot], cudaEventDisableTiming));
  CK(cudaEventRecord(E->fev[slot], E->stream));
}

// the host's view of a slot's flags is complete
static void wait_flags(dem_engine *E, int slot)
{
  if (E->fev_peer[slot]) {  // written by k_wait into the page-locked block; the serial word arrives last
    const volatile int *ser = (const volatile int *)E->hflag + 16 + slot;
    const auto t0 = std::chrono::steady_clock::now();
    long spins = 0;
    while (*ser != E->fser_slot[slot]) {
      if ((++spins & 0xfff) == 0) {
        if (cudaStreamQuery(E->stream) == cudaSuccess && *ser != E->fser_slot[slot]) dem_fail(E, DEM_ERR_CUDA, "step flags did not arrive (stream idle)");
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) { E->comm_bad = 1; dem_fail(E, DEM_ERR_CUDA, "timed out waiting for the step flags of the other ranks"); }
      }
    }
    return;
  }
  if (E->fev[slot]) CK(cudaEventSynchronize(E->fev[slot]));
}

// fused ghost push: usable for this launch?  (every step kernel of the full list ends in step_epilogue, which stores the copies;
// the owner-list option keeps the pack-kernel path)
static bool fused_on(dem_engine *E, int mode)
{
  return E->fz_ready && mode != MODE_SETUP && E->ls[E->lcur].fmt == 0 && E->nlocal > 0;
}
// One step on the stream: the step launch, the ghost refresh, the flag hand-over.
//  * fused ghost push (fz_ready): the step kernel stores every copy of a particle's new records from its epilogue; several
//    ranks: one k_handshake launch publishes the exchange + this rank's flags and waits for the neighbours' / everybody's;
//    one rank: one k_flags_host launch hands the flags to the host.
//  * otherwise: pack kernels / k_push + k_wait (forward_comm) and post_flags.
static void step_and_comm(dem_engine *E, int mode, int slot)
{
  const bool fz = fused_on(E, mode);
  E->fz_on = fz ? 1 : 0;
  launch_step(E, mode, true);
  E->fz_on = 0;
  E->cur ^= 1;
  if (!fz) { const bool sent = forward_comm(E, slot); post_flags(E, slot, sent); E->slot_zeroed = sent ? 1 : 0; return; }
  E->fserial++; E->fser_slot[slot] = E->fserial;
  if (E->nranks == 1) {
    if (!E->hflag_dev) { void *dp = nullptr; CK(cudaHostGetDevicePointer(&dp, E->hflag, 0)); E->hflag_dev = (int *)dp; }
    k_flags_host<<<1, 32, 0, E->stream>>>(flag_slot(E, slot), E->hflag_dev + 4 * slot, E->hflag_dev + 16 + slot, E->fserial, flag_slot(E, slot ^ 1));
  } else {
    ShakeP Q; memset(&Q, 0, sizeof Q);
    WaitP &Wt = Q.W;
    for (int sd = 0; sd < 2; sd++) {
      dem_engine::Swap &W = E->swaps[E->fz_rq[sd]];
      W.serial++;
      const int peer_send = neighbor_rank(E, W.dim, W.side ? 1 : -1), peer_recv = neighbor_rank(E, W.dim, W.side ? -1 : 1);
      if (peer_send >= 0) { Q.sig_out[sd] = W.psig + E->fz_rq[sd]; Q.serial_out[sd] = W.serial; }
      if (peer_recv >= 0) { Wt.sig[sd] = E->hsig.p + E->fz_rq[sd]; Wt.serial[sd] = W.serial; }
    }
    Q.myflags = flag_slot(E, slot); Q.me = E->rank;
    for (int r = 0; r < E->nranks; r++) Q.peer_box[r] = E->peer_fbox[r];
    Wt.with_flags = 1; Wt.box = E->fbox.p; Wt.nranks = E->nranks; Wt.slot = slot; Wt.fserial = E->fserial;
    Wt.gate_out = flag_slot(E, slot) + 4; Wt.host_out = E->hflag_dev + 4 * slot; Wt.host_serial = E->hflag_dev + 16 + slot;
    Wt.timeout_flag = E->hsig.p + 15; Wt.zero_next = flag_slot(E, slot ^ 1);
    k_handshake<<<1, 32, 0, E->stream>>>(Q);
  }
  E->launches++;
  E->fev_peer[slot] = 1;  // wait_flags spins on the serial word of the page-locked block
  E->slot_zeroed = 1;
}

static void collect_timing(dem_engine *E)
{
  for (long k = 0; k < E->ev_used; k++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, E->ev[2 * k], E->ev[2 * k + 1]) == cudaSuccess) { E->step_ms += ms; E->step_calls++; }
  }
  E->ev_used = 0;
}

// what Verlet::setup needs before the lists are built: tables, cutoffs, cell grid, triangle grid, wall history rows
static void setup_prepare(dem_engine *e)
{
  dem_engine *E = e;  // (CK / NK / dem_fail name the engine E)
  if (!e->uploaded) dem_fail(e, DEM_ERR_STATE, "setup before dem_upload_particles");
  if (!(e->dt > 0)) dem_fail(e, DEM_ERR_STATE, "timestep not set");
  if (!e->have_pair && e->walls.empty() && e->mwalls.empty()) dem_fail(e, DEM_ERR_STATE, "no pair_style and no wall defined");
  CK(cudaSetDevice(e->device));
  if (!e->setup_done || e->dirty) {
    // (also after the first run: `neighbor`, `fix property/global` or the box changed between two runs -- tables, cutoff, cell
    // grid and the triangle grid are derived again; lists are rebuilt below anyway)
    const bool first = !e->setup_done;
    if (e->nranks > 1) {
      // every rank may have been handed only its own share of the particles (or none at all): the neighbour cutoff and the
      // bond models' contact-distance factor need the GLOBAL extreme radii (the reference: MPI_Allreduce in
      // pair_gran.cpp:591-603 / neighbor.cpp)
      e->cnt_dev.ensure(e, 64);
      double h[2] = {e->rmax, e->rmin > 0.0 ? -e->rmin : -1e300};
      double *d = (double *)(e->cnt_dev.p + 16);
      CK(cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, e->stream));
      NK(g_nccl.AllReduce(d, d + 2, 2, ncclDouble, ncclMax, e->comm, e->stream));
      CK(cudaMemcpyAsync(h, d + 2, sizeof h, cudaMemcpyDeviceToHost, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      e->rmax = h[0]; e->rmin = -h[1];
    }
    derive_tables(e);
    setup_grid(e);
    if (!first) e->grid_ready = 0;
    mesh_prepare(e);
    e->dirty = 0;
    if (first && e->nwrows) {
      e->whist.release(); e->whist.ensure(e, (size_t)e->nwrows * e->cap);
      CK(cudaMemsetAsync(e->whist.p, 0, (size_t)e->nwrows * e->cap * sizeof(double), e->stream));
      e->whist_tmp.release(); e->whist_tmp.ensure(e, (size_t)e->nwrows * e->cap);
    }
  }
  if (e->have_pair && e->pm.cohesion == C_BOND && !e->bondc.p) { e->bondc.ensure(e, 4); CK(cudaMemsetAsync(e->bondc.p, 0, 4 * sizeof(unsigned long long), e->stream)); }
}

extern "C" int dem_setup(dem_engine *e)
{
  API_BEGIN
  if (e->ins_open) dem_fail(e, DEM_ERR_STATE, "dem_setup inside an insertion step");
  setup_prepare(e);
  clear_flags(e);
  rebuild(e);
  e->nbuilds = 0;  // neighbor->ncalls counts the builds of the current run only
  launch_step(e, MODE_SETUP, false);
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaGetLastError());
  e->setup_done = 1; e->forces_valid = 1; e->need_setup = 0;
  API_END
}

extern "C" int dem_run(dem_engine *e, long nsteps)
{
  API_BEGIN
  if (!e->setup_done) dem_fail(e, DEM_ERR_STATE, "dem_run before dem_setup");
  if (e->need_setup) dem_fail(e, DEM_ERR_STATE, "particles were inserted: dem_setup before dem_run");
  if (e->ins_open) dem_fail(e, DEM_ERR_STATE, "dem_run inside an insertion step");
  if (nsteps < 0) dem_fail(e, DEM_ERR_ARG, "nsteps < 0");
  if (nsteps == 0) return DEM_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  e->step_ms = 0; e->step_calls = 0; e->ev_used = 0;
  // first half step from the stored forces (fix_nve_sphere.cpp:134-183)
  clear_flags(e);
  e->fslot = 0; e->gate = nullptr; e->gate_mask = 0;
  if (e->nlocal) {
    StepP P = step_params(e, MODE_STEP);
    k_initial_integrate<<<GRID(P.nlocal, 256), 256, 0, st>>>(P);
    e->launches++;
  }
  e->cur ^= 1;
  post_flags(e, 0, forward_comm(e, 0));
  const bool moving = e->any_moving && e->mesh_ready;
  int overflow_seen = 0;
  // one step on the device: mesh motion (fix move/mesh initial_integrate, fix_move_mesh.cpp:221-238), the fused step
  // kernel(s), the ghost refresh and the flag hand-over
  auto issue_step = [&](long s, int slot) {
    e->fslot = slot;
    if (!e->slot_zeroed) CK(cudaMemsetAsync(flag_slot(e, slot), 0, 8 * sizeof(int), st));  // (else the previous hand-over zeroed it: k_wait / the step kernel's last block)
    if (moving) {
      MeshP M = mesh_params(e);
      for (size_t m = 0; m < e->meshes.size(); m++)
        if (e->meshes[m].moving) { mesh_launch_move(M, (int)m, e->dt, 0.25 * e->skin * e->skin, flag_slot(e, slot), e->gate, e->gate_mask, st); e->launches++; }
    }
    step_and_comm(e, s == nsteps ? MODE_LAST : MODE_STEP, slot);
  };
  for (long s = 1; s <= nsteps; s++) {
    e->ntimestep++;
    const int prev = (int)((s - 1) & 1), slot = (int)(s & 1);
    // Neighbor::decide (neighbor.cpp:1362-1376): a fix may force the rebuild (fix->next_reneighbor = the mesh trigger of
    // the previous step), else the distance check when it is due
    const int ago1 = e->ago + 1;
    const bool due = ago1 >= e->delay && ago1 % e->every == 0;
    const int mask = ((due && e->check) ? 1 : 0) | (moving ? 4 : 0);
    bool rebuild_now = due && !e->check;
    if (!rebuild_now) {
      // speculative: queue the step gated on the previous step's flags, then look at those flags
      e->gate = mask ? gate_slot(e, prev) : nullptr; e->gate_mask = mask;
      const long ev0 = e->ev_used, l0 = e->launches;
      issue_step(s, slot);
      wait_flags(e, prev);
      const int *hf = e->hflag + 4 * prev;
      overflow_seen |= hf[1];
      if (((mask & 1) && hf[0]) || ((mask & 4) && hf[2])) {  // the queued step returned at its gate: undo the host side
        rebuild_now = true;
        e->cur ^= 1; e->ev_used = ev0; e->launches = l0;
      } else e->ago = ago1;
    }
    if (rebuild_now) {
      // (forced rebuild, `check no`: the previous step's flags were not looked at above -- its history-slot overflow must
      // not be wiped by clear_flags)
      wait_flags(e, prev); overflow_seen |= e->hflag[4 * prev + 1];
      e->gate = nullptr; e->gate_mask = 0;
      if (moving) {  // the mesh moves before the lists are rebuilt
        MeshP M = mesh_params(e);
        for (size_t m = 0; m < e->meshes.size(); m++)
          if (e->meshes[m].moving) { mesh_launch_move(M, (int)m, e->dt, 0.25 * e->skin * e->skin, flag_slot(e, slot), nullptr, 0, st); e->launches++; }
      }
      rebuild(e);
      clear_flags(e);
      const bool was_moving = false;
      (void)was_moving;
      // the step itself, not gated; its mesh motion has been done above
      e->fslot = slot;
      step_and_comm(e, s == nsteps ? MODE_LAST : MODE_STEP, slot);
    }
    if (e->ev_used >= 2048) { CK(cudaStreamSynchronize(st)); collect_timing(e); }
  }
  e->gate = nullptr; e->gate_mask = 0;
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  collect_timing(e);
  if (e->nranks > 1 && e->p2p_ok) {
    int to = 0;
    CK(cudaMemcpy(&to, e->hsig.p + 15, sizeof(int), cudaMemcpyDeviceToHost));
    if (to) { e->comm_bad = 1; dem_fail(e, DEM_ERR_CUDA, "halo exchange timed out waiting for a neighbour rank"); }
  }
  if (overflow_seen || e->hflag[1] || e->hflag[5]) { e->hflag[1] = e->hflag[5] = 0; dem_fail(e, DEM_ERR_OVERFLOW, "a particle gained more new contacts between two rebuilds than free history slots; raise option 'histslots'"); }
  e->forces_valid = 1;
  API_END
}

// The timestep in which fix insert/* creates particles, in two halves (FixInsert::pre_exchange, fix_insert.cpp:672-905, runs
// between the first half step and the forced rebuild of that timestep, verlet.cpp:277-310).  Between the two calls the
// caller sees the positions the reference's overlap check sees (dem_download "x").
//   begin: first half step of the particles the engine holds (nothing to do while it holds none)
//   end:   the new particles appear (mass = density * volume as FixTemplateSphere forms it), lists are rebuilt, forces
//          evaluated, second half step for everybody; n == 0 is a step with a forced rebuild and no newcomer
extern "C" int dem_insert_step_begin(dem_engine *e)
{
  API_BEGIN
  if (e->ins_open) dem_fail(e, DEM_ERR_STATE, "dem_insert_step_begin called twice");
  if (e->uploaded) {
    if (!e->setup_done) dem_fail(e, DEM_ERR_STATE, "dem_insert_step_begin before dem_setup");
    if (e->need_setup) dem_fail(e, DEM_ERR_STATE, "particles were inserted: dem_setup before dem_insert_step_begin");
    CK(cudaSetDevice(e->device));
    clear_flags(e);
    e->fslot = 0; e->gate = nullptr; e->gate_mask = 0;
    if (e->nlocal) {
      StepP P = step_params(e, MODE_STEP);
      k_initial_integrate<<<GRID(P.nlocal, 256), 256, 0, e->stream>>>(P);
      e->launches++;
    }
    e->cur ^= 1;
    CK(cudaStreamSynchronize(e->stream));
  }
  e->ins_open = 1;
  API_END
}

extern "C" int dem_insert_step_end(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x,
                                   const double *v, const double *omega, const double *radius, const double *density)
{
  if (!e) return DEM_ERR_ARG;
  if (!e->ins_open) { e->err = "dem_insert_step_end without dem_insert_step_begin"; return DEM_ERR_STATE; }
  if (!e->uploaded && n <= 0) { e->ins_open = 0; e->ntimestep++; return DEM_OK; }  // a step of an empty box
  e->ins_mass = 1;
  const int rc = n > 0 ? dem_insert_particles(e, n, tag, type, mask, x, v, omega, radius, density) : DEM_OK;
  e->ins_mass = 0;
  if (rc != DEM_OK) return rc;
  API_BEGIN
  CK(cudaSetDevice(e->device));
  if (!(e->dt > 0)) dem_fail(e, DEM_ERR_STATE, "timestep not set");
  if (!e->have_pair && e->walls.empty() && e->mwalls.empty()) dem_fail(e, DEM_ERR_STATE, "no pair_style and no wall defined");
  const bool first = !e->setup_done;
  setup_prepare(e);
  cudaStream_t st = e->stream;
  e->step_ms = 0; e->step_calls = 0; e->ev_used = 0;
  e->ntimestep++;
  clear_flags(e);
  e->fslot = 1; e->gate = nullptr; e->gate_mask = 0;
  if (e->any_moving && e->mesh_ready && !first) {  // the mesh moves before the lists are rebuilt (dem_run)
    MeshP M = mesh_params(e);
    for (size_t m = 0; m < e->meshes.size(); m++)
      if (e->meshes[m].moving) { mesh_launch_move(M, (int)m, e->dt, 0.25 * e->skin * e->skin, flag_slot(e, 1), nullptr, 0, st); e->launches++; }
  }
  if (first) e->nbuilds = 0;  // (the run of an empty box had no setup that would have reset the count of builds)
  rebuild(e);
  clear_flags(e);
  e->fslot = 1;
  step_and_comm(e, MODE_LAST, 1);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  collect_timing(e);
  e->ins_open = 0;
  if (e->hflag[1] || e->hflag[5]) { e->hflag[1] = e->hflag[5] = 0; dem_fail(e, DEM_ERR_OVERFLOW, "a particle gained more new contacts between two rebuilds than free history slots; raise option 'histslots'"); }
  e->setup_done = 1; e->forces_valid = 1; e->need_setup = 0;
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

// ascending-tag permutation on the device (radix sort of the tags), cached until the next rebuild
static void ensure_order(dem_engine *E)
{
  if (E->order_valid) return;
  const int n = (int)E->nlocal;
  cudaStream_t st = E->stream;
  if (n) {
    E->order.ensure(E, E->cap); E->order_keys.ensure(E, E->cap); E->vals.ensure(E, E->cap); E->keys.ensure(E, E->cap);
    ensure_cub(E, (size_t)E->cap);
    k_iota<<<GRID(n, 256), 256, 0, st>>>(n, E->vals.p);
    size_t tb = E->cubtmp.n;
    CK(cub::DeviceRadixSort::SortPairs(E->cubtmp.p, tb, (const unsigned *)E->tag.p, (unsigned *)E->order_keys.p, E->vals.p, E->order.p, n, 0, 32, st));
    E->launches += 2;
  }
  E->order_valid = 1;
}

extern "C" int dem_download(dem_engine *e, const char *field, void *out, long count)
{
  API_BEGIN
  if (!e->uploaded) dem_fail(e, DEM_ERR_STATE, "no particles");
  if (count != e->nlocal) dem_fail(e, DEM_ERR_ARG, "count %ld != nlocal %ld", count, e->nlocal);
  CK(cudaSetDevice(e->device));
  const int n = (int)e->nlocal;
  if (!n) return DEM_OK;
  cudaStream_t st = e->stream;
  ensure_order(e);
  std::string f(field);
  const int c = e->cur;
  e->stage.ensure(e, (size_t)n * 3 * sizeof(double) + 64);
  double *od = (double *)e->stage.p; int *oi = (int *)e->stage.p;
  size_t bytes = 0;
  if (f == "tag") { CK(cudaMemcpyAsync(out, e->order_keys.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); return DEM_OK; }
  else if (f == "type" || f == "mask") { k_gather_out<<<GRID(n, 256), 256, 0, st>>>(n, e->order.p, e->wt[c].p, f == "type" ? 4 : 5, od, oi); bytes = (size_t)n * sizeof(int); }
  else if (f == "radius") { k_gather_out<<<GRID(n, 256), 256, 0, st>>>(n, e->order.p, e->xr[c].p, 3, od, oi); bytes = (size_t)n * sizeof(double); }
  else if (f == "rmass") { k_gather_out<<<GRID(n, 256), 256, 0, st>>>(n, e->order.p, e->vm[c].p, 3, od, oi); bytes = (size_t)n * sizeof(double); }
  else if (f == "density") { k_gather_rows_out<<<GRID(n, 256), 256, 0, st>>>(n, e->order.p, e->density.p, 0, 1, od); bytes = (size_t)n * sizeof(double); }
  else if (f == "x" || f == "v" || f == "omega") {
    k_gather_out<<<GRID(n, 256), 256, 0, st>>>(n, e->order.p, f == "x" ? e->xr[c].p : f == "v" ? e->vm[c].p : e->wt[c].p, 0, od, oi);
    bytes = (size_t)n * 3 * sizeof(double);
  } else if (f == "f" || f == "torque") {
    if (!e->forces_valid) dem_fail(e, DEM_ERR_STATE, "forces are only available after setup or run");
    k_gather_rows_out<<<GRID(n, 256), 256, 0, st>>>(n, e->order.p, f == "f" ? e->f.p : e->tq.p, (size_t)e->cap, 3, od);
    bytes = (size_t)n * 3 * sizeof(double);
  } else dem_fail(e, DEM_ERR_ARG, "unknown field %s", field);
  e->launches++;
  CK(cudaMemcpyAsync(out, e->stage.p, bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  API_END
}

struct PairRow { int lo, hi, flag; long src; int has; };
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
  std::vector<int> nown(n, 0);
  if (L.fmt) for (long i = 0; i < n; i++) { nown[i] = NN2_OWN(nn[i]); nn[i] = NN2_TOT(nn[i]); }
  else for (long i = 0; i < n; i++) nn[i] &= 0xffff;
  int kmax = 0; for (long i = 0; i < n; i++) kmax = std::max(kmax, nn[i]);
  if (L.fmt) kmax = L.maxk;  // (owner list: partner-owned entries sit at the back of the row)
  std::vector<unsigned> nbr((size_t)kmax * L.cap); std::vector<int> ptag((size_t)kmax * L.cap);
  if (kmax) {
    CK(cudaMemcpy(nbr.data(), L.nbr.p, nbr.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ptag.data(), L.ptag.p, ptag.size() * sizeof(int), cudaMemcpyDeviceToHost));
  }
  hist.assign((size_t)L.hslots * dnum * L.cap, make_double4(0., 0., 0., 0.));
  if (kmax && dnum) CK(cudaMemcpy(hist.data(), L.hist.p, hist.size() * sizeof(double4), cudaMemcpyDeviceToHost));
  const ModelP &M = E->pm;
  auto comp = [](const double4 &v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; };
  if (L.fmt) {
    // owner list: one row per OWNED entry (a pair across a periodic face or a brick boundary is owned on both sides: the
    // reference lists such a pair on both sides, too); history is stored "lower tag first" whoever owns the pair
    for (long i = 0; i < n; i++) for (int k = 0; k < nown[i]; k++) {
      const unsigned w = nbr[(size_t)k * L.cap + i];
      const int tj = ptag[(size_t)k * L.cap + i];
      const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
      rows.push_back(PairRow{std::min(tags[i], tj), std::max(tags[i], tj), slot >= 0 ? 1 : 0, (long)std::max(slot, 0) * L.cap + i, slot >= 0 ? 1 : 0});
    }
  } else
  for (long i = 0; i < n; i++) for (int k = 0; k < nn[i]; k++) {
    const unsigned w = nbr[(size_t)k * L.cap + i];
    const int tj = ptag[(size_t)k * L.cap + i];
    const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
    int flag = slot >= 0 ? 1 : 0;
    if (flag && E->have_pair && M.cohesion) {  // reference contact_flags != 0 <=> bonded or sticky (see k_step_bond)
      const double b0 = hist[(size_t)(slot * dnum + M.rec_bond) * L.cap + i].x;
      const double S = comp(hist[(size_t)(slot * dnum + M.rec_bond + M.nbond / 4) * L.cap + i], M.nbond % 4);
      flag = (b0 != 0.0 || S != 0.0) ? 1 : 0;
    }
    if (tags[i] < tj) rows.push_back(PairRow{tags[i], tj, flag, (long)std::max(slot, 0) * L.cap + i, slot >= 0 ? 1 : 0});
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
      if (rows[r].has) {
        if (M.cohesion) for (int d = 0; d < M.nbond; d++) {
          const double4 v = h[(size_t)(k * nrec + M.rec_bond + d / 4) * L.cap + i];
          hist[r * dn + M.off_bond + d] = (d % 4 == 0) ? v.x : (d % 4 == 1) ? v.y : (d % 4 == 2) ? v.z : v.w;
        }
        if (M.rec_norm >= 0) for (int d = 0; d < 12; d++) {
          const double4 v = h[(size_t)(k * nrec + M.rec_norm + d / 4) * L.cap + i];
          hist[r * dn + M.off_norm + d] = (d % 4 == 0) ? v.x : (d % 4 == 1) ? v.y : (d % 4 == 2) ? v.z : v.w;
        }
        if (M.rec_shear >= 0) { const double4 v = h[(size_t)(k * nrec + M.rec_shear) * L.cap + i]; hist[r * dn + M.off_shear] = v.x; hist[r * dn + M.off_shear + 1] = v.y; hist[r * dn + M.off_shear + 2] = v.z;
          if (M.tangential == 2) hist[r * dn + M.off_shear + 3] = v.w; }
        if (M.rec_roll >= 0) { const double4 v = h[(size_t)(k * nrec + M.rec_roll) * L.cap + i]; hist[r * dn + M.off_roll] = v.x; hist[r * dn + M.off_roll + 1] = v.y; hist[r * dn + M.off_roll + 2] = v.z; }
      }
    }
  }
  API_END
}

// ---- per-contact output (option contact_output): rows (own tag, partner tag, force on me, torque on me) of the last force
// evaluation that materialised forces -- dem_setup or the last step of dem_run --, ordered by (own tag, partner tag)
struct ContactRows { std::vector<int> tags; std::vector<double> v; };
static void collect_contacts(dem_engine *E, ContactRows &R)
{
  R.tags.clear(); R.v.clear();
  ListSet &L = E->ls[E->lcur];
  if (!E->have_pair || E->pm.cohesion || L.fmt != 0) dem_fail(E, DEM_ERR_UNSUPPORTED, "per-contact output covers the plain contact models on the default (full) list");
  if (!(E->opt.count("contact_output") && E->opt["contact_output"] != 0)) dem_fail(E, DEM_ERR_STATE, "per-contact output needs option contact_output 1 (before dem_setup)");
  if (!E->forces_valid || !L.valid || !E->cout.p) dem_fail(E, DEM_ERR_STATE, "per-contact output is available after dem_setup or dem_run");
  const int n = (int)E->nlocal;
  if (!n) return;
  cudaStream_t st = E->stream;
  StepP P = step_params(E, MODE_SETUP);
  P.cout = E->cout.p; P.serial = (double)E->cout_serial;
  E->flo.ensure(E, n + 1); E->slo.ensure(E, n + 1);
  ensure_cub(E, (size_t)E->cap);
  k_contact_count<<<GRID(n, 256), 256, 0, st>>>(P, E->flo.p);
  size_t tb = E->cubtmp.n;
  CK(cub::DeviceScan::ExclusiveSum(E->cubtmp.p, tb, E->flo.p, E->slo.p, n, st));
  int last[2];
  CK(cudaMemcpyAsync(&last[0], E->flo.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&last[1], E->slo.p + n - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const size_t rows = (size_t)last[0] + last[1];
  E->launches += 2;
  if (!rows) return;
  DevBuf<int> dt; DevBuf<double> dv;
  dt.ensure(E, 2 * rows); dv.ensure(E, 6 * rows);
  k_contact_fill<<<GRID(n, 256), 256, 0, st>>>(P, L.ptag.p, E->tag.p, E->slo.p, dt.p, dv.p);
  E->launches++;
  std::vector<int> ht(2 * rows); std::vector<double> hv(6 * rows);
  CK(cudaMemcpyAsync(ht.data(), dt.p, ht.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hv.data(), dv.p, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  dt.release(); dv.release();
  std::vector<size_t> o(rows);
  for (size_t r = 0; r < rows; r++) o[r] = r;
  std::sort(o.begin(), o.end(), [&](size_t a, size_t b) { return ht[2 * a] != ht[2 * b] ? ht[2 * a] < ht[2 * b] : ht[2 * a + 1] < ht[2 * b + 1]; });
  R.tags.resize(2 * rows); R.v.resize(6 * rows);
  for (size_t r = 0; r < rows; r++) { R.tags[2 * r] = ht[2 * o[r]]; R.tags[2 * r + 1] = ht[2 * o[r] + 1]; memcpy(&R.v[6 * r], &hv[6 * o[r]], 6 * sizeof(double)); }
}
// compute bond/counter (compute_bond_counter.cpp:101-138), see include/dem_b200.h.  Only the linear bond model feeds it in the
// reference: bond/nonlinear looks for a compute style "bond/nonlinear/counter" that does not exist
// (cohesion_model_bond_nonlinear.h:381), so its counter stays at zero -- reproduced here.
extern "C" int dem_bond_counter(dem_engine *e, double *out6)
{
  API_BEGIN
  if (!out6) dem_fail(e, DEM_ERR_ARG, "null output");
  for (int k = 0; k < 6; k++) out6[k] = 0.0;
  if (!e->setup_done) dem_fail(e, DEM_ERR_STATE, "dem_bond_counter before dem_setup");
  if (!(e->have_pair && e->pm.cohesion == C_BOND)) return DEM_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  unsigned long long h[2] = {0, 0};
  if (e->nranks > 1) NK(g_nccl.AllReduce(e->bondc.p, e->bondc.p, 2, ncclUint64, ncclSum, e->comm, st));
  CK(cudaMemcpyAsync(h, e->bondc.p, sizeof h, cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(e->bondc.p, 0, 2 * sizeof(unsigned long long), st));
  CK(cudaStreamSynchronize(st));
  out6[0] = (double)h[0]; out6[1] = (double)h[1];
  out6[2] = (double)((unsigned int)h[0] - (unsigned int)h[1]);  // counted (0 between runs) + created - broken, unsigned like the reference
  API_END
}

extern "C" int dem_contact_count(dem_engine *e, long *n)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  ContactRows R; collect_contacts(e, R);
  if (n) *n = (long)(R.tags.size() / 2);
  API_END
}
extern "C" int dem_download_contacts(dem_engine *e, int *tag, int *partner, double *force, double *torque)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  ContactRows R; collect_contacts(e, R);
  const size_t rows = R.tags.size() / 2;
  for (size_t r = 0; r < rows; r++) {
    if (tag) tag[r] = R.tags[2 * r];
    if (partner) partner[r] = R.tags[2 * r + 1];
    for (int d = 0; d < 3; d++) {
      if (force) force[3 * r + d] = R.v[6 * r + d];
      if (torque) torque[3 * r + d] = R.v[6 * r + 3 + d];
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

// ---- mesh read-back --------------------------------------------------------------------------
extern "C" int dem_download_mesh(dem_engine *e, const char *mesh_id, const char *field, void *out, long count)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  for (auto &m : e->meshes) if (m.id == mesh_id) {
    const long T = m.ntri;
    std::string f(field);
    if (f == "nodes") {
      if (count != 9 * T) dem_fail(e, DEM_ERR_ARG, "nodes: count must be 9*ntri");
      if (e->mesh_ready) {
        std::vector<TriRec> h(T);
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaMemcpy(h.data(), e->dtri.p + m.first, T * sizeof(TriRec), cudaMemcpyDeviceToHost));
        for (long t = 0; t < T; t++) memcpy((double *)out + 9 * t, h[t].node, 9 * sizeof(double));
      } else memcpy(out, m.nodes.data(), 9 * T * sizeof(double));
      return DEM_OK;
    }
    if (!e->mesh_ready) dem_fail(e, DEM_ERR_STATE, "mesh topology is available after setup");
    if (f == "edge_vec" || f == "edge_norm" || f == "surf_norm" || f == "center") {
      const int w = (f == "edge_vec" || f == "edge_norm") ? 9 : 3;
      if (count != w * T) dem_fail(e, DEM_ERR_ARG, "%s: count must be %d*ntri", field, w);
      std::vector<TriRec> h(T);
      CK(cudaStreamSynchronize(e->stream));
      CK(cudaMemcpy(h.data(), e->dtri.p + m.first, T * sizeof(TriRec), cudaMemcpyDeviceToHost));
      for (long t = 0; t < T; t++)
        memcpy((double *)out + w * t, f == "edge_vec" ? h[t].edgeVec : f == "edge_norm" ? h[t].edgeNorm : f == "surf_norm" ? h[t].surfNorm : h[t].center, w * sizeof(double));
      return DEM_OK;
    }
    const std::vector<int> *src = f == "edge_active" ? &m.edge_active : f == "corner_active" ? &m.corner_active : f == "obtuse" ? &m.obtuse : f == "nneighs" ? &m.nneighs : nullptr;
    if (!src) dem_fail(e, DEM_ERR_ARG, "unknown mesh field %s", field);
    if (count != (long)src->size()) dem_fail(e, DEM_ERR_ARG, "%s: count %ld != %ld", field, count, (long)src->size());
    memcpy(out, src->data(), src->size() * sizeof(int));
    return DEM_OK;
  }
  dem_fail(e, DEM_ERR_ARG, "no mesh with id %s", mesh_id);
  API_END
}

struct MeshRow { int tag, tri; long src; };
static void collect_mesh_rows(dem_engine *E, const MeshHost &m, std::vector<MeshRow> &rows, std::vector<double4> &hist)
{
  rows.clear();
  if (!E->mesh_ready || !E->nlocal) return;
  CK(cudaStreamSynchronize(E->stream));
  const long n = E->nlocal;
  std::vector<int> tags(n), mi((size_t)(1 + E->mslots) * E->cap);
  CK(cudaMemcpy(tags.data(), E->tag.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(mi.data(), E->mint[E->mcur].p, mi.size() * sizeof(int), cudaMemcpyDeviceToHost));
  hist.assign((size_t)E->mslots * E->mhrec * E->cap, make_double4(0., 0., 0., 0.));
  if (E->mhrec) CK(cudaMemcpy(hist.data(), E->mhist[E->mcur].p, hist.size() * sizeof(double4), cudaMemcpyDeviceToHost));
  for (long i = 0; i < n; i++) for (int s = 0; s < E->mslots; s++) {
    const int t = mi[(size_t)(1 + s) * E->cap + i];
    if (t >= m.first && t < m.first + m.ntri) rows.push_back(MeshRow{tags[i], t - m.first, (long)s * E->cap + i});
  }
  std::sort(rows.begin(), rows.end(), [](const MeshRow &a, const MeshRow &b) { return a.tag != b.tag ? a.tag < b.tag : a.tri < b.tri; });
}
extern "C" int dem_mesh_force(dem_engine *e, const char *mesh_id, double *out9)
{
  API_BEGIN
  if (!out9) dem_fail(e, DEM_ERR_ARG, "null output");
  CK(cudaSetDevice(e->device));
  for (size_t m = 0; m < e->meshes.size(); m++) if (e->meshes[m].id == mesh_id) {
    if (!e->meshes[m].stress) dem_fail(e, DEM_ERR_STATE, "mesh %s does not track stress (fix mesh/surface/stress)", mesh_id);
    if (!e->mesh_ready) { for (int d = 0; d < 6; d++) out9[d] = 0.0; for (int d = 0; d < 3; d++) out9[6 + d] = e->meshes[m].p_ref[d]; return DEM_OK; }
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out9, e->dmforce.p + 6 * m, 6 * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out9 + 6, e->dmpref.p + 3 * m, 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return DEM_OK;
  }
  dem_fail(e, DEM_ERR_ARG, "no mesh with id %s", mesh_id);
  API_END
}
extern "C" int dem_mesh_contact_count(dem_engine *e, const char *mesh_id, long *n, int *dnum)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  for (auto &m : e->meshes) if (m.id == mesh_id) {
    std::vector<MeshRow> rows; std::vector<double4> h;
    collect_mesh_rows(e, m, rows, h);
    if (n) *n = (long)rows.size();
    if (dnum) *dnum = m.wall >= 0 ? e->mwalls[m.wall].m.dnum : 0;
    return DEM_OK;
  }
  dem_fail(e, DEM_ERR_ARG, "no mesh with id %s", mesh_id);
  API_END
}
extern "C" int dem_download_mesh_contacts(dem_engine *e, const char *mesh_id, int *tag, int *tri, double *hist)
{
  API_BEGIN
  CK(cudaSetDevice(e->device));
  for (auto &m : e->meshes) if (m.id == mesh_id) {
    std::vector<MeshRow> rows; std::vector<double4> h;
    collect_mesh_rows(e, m, rows, h);
    const ModelP *W = m.wall >= 0 ? &e->mwalls[m.wall].m : nullptr;
    const int dn = W ? W->dnum : 0;
    for (size_t r = 0; r < rows.size(); r++) {
      if (tag) tag[r] = rows[r].tag;
      if (tri) tri[r] = rows[r].tri;
      if (hist && W) {
        const long s = rows[r].src / e->cap, i = rows[r].src % e->cap;
        for (int d = 0; d < dn; d++) hist[r * dn + d] = 0.0;
        if (W->rec_shear >= 0) { const double4 v = h[(size_t)(s * e->mhrec + W->rec_shear) * e->cap + i]; hist[r * dn + W->off_shear] = v.x; hist[r * dn + W->off_shear + 1] = v.y; hist[r * dn + W->off_shear + 2] = v.z; }
        if (W->rec_roll >= 0) { const double4 v = h[(size_t)(s * e->mhrec + W->rec_roll) * e->cap + i]; hist[r * dn + W->off_roll] = v.x; hist[r * dn + W->off_roll + 1] = v.y; hist[r * dn + W->off_roll + 2] = v.z; }
      }
    }
    return DEM_OK;
  }
  dem_fail(e, DEM_ERR_ARG, "no mesh with id %s", mesh_id);
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
    if (L.fmt) k_count_pairs2<<<GRID(e->nlocal, 256), 256, 0, e->stream>>>((int)e->nlocal, L.numneigh.p, L.nbr.p, L.cap, L.maxk, e->counters.p);
    else k_count_pairs<<<GRID(e->nlocal, 256), 256, 0, e->stream>>>((int)e->nlocal, L.numneigh.p, L.nbr.p, L.cap, e->counters.p);
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, e->counters.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    s->npairs_full = (long)h[0]; s->ncontacts_full = (long)h[1];
  }
  s->step_kernel_ms = e->step_ms; s->step_kernel_calls = e->step_calls;
  API_END
}
