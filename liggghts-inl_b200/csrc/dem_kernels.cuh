// dem_kernels.cuh -- hand-written sm_100a kernels of the DEM timestep engine.
//
// Design (B200-first, not a port of the reference loops):
//  * 32-byte particle records (x|r, v|m, omega|type+mask) so that one neighbour gather is
//    exactly one DRAM sector; records are double-buffered so that the whole timestep
//    (pair forces + walls + gravity + freeze + final_integrate(n) + initial_integrate(n+1) +
//    rebuild trigger) is ONE kernel with no force array round trip through HBM.
//  * FULL neighbour list in transposed ELLPACK layout, every pair evaluated by both owners
//    in one canonical orientation (lower tag first) -> no atomics, bit-reproducible runs,
//    and both copies of a pair's history stay bit-identical.
//  * contact history is valid only when the NBR_HIST bit of the neighbour word is set
//    (== the reference's contact_flag != 0), so untouched pairs cost no history traffic.
// Reference behaviour followed: see dem_contact.cuh and the per-kernel notes below.
#pragma once
#include "dem_contact.cuh"

namespace dem {

__device__ __forceinline__ double sq3_rn(double a, double b, double c)
{  // a*a+b*b+c*c exactly as the un-contracted CPU expression (predicates must be bit-exact)
  return __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c));
}
// 256-bit record access (sm_100: LDG.E.256 / STG.E.256): one instruction per 32-byte particle record
__device__ __forceinline__ double4 ldg4(const double4 *p)
{
  double4 v;
  // not volatile: a read-only load the compiler may hoist and batch with its neighbours
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
#ifndef DEM_STREAM_LD
#define DEM_STREAM_LD 1   // read-once data (own records, neighbour words, history records) bypasses L1: 0.9108 -> 0.9057 ms
#endif
__device__ __forceinline__ double4 ldg4s(const double4 *p)
{  // read-once record (own particle): no L1 allocation when DEM_STREAM_LD
  double4 v;
#if DEM_STREAM_LD
  asm("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#else
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#endif
  return v;
}
__device__ __forceinline__ void st4(double4 *p, const double4 &v)
{
  // volatile (must not be dropped) but no memory clobber: nothing in the same thread reads it back
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w));
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ bool step_gated(const StepP &P)
{  // see StepP::gate
  return P.gate && (((P.gate_mask & 1) && P.gate[0]) || ((P.gate_mask & 4) && P.gate[2]));
}
__device__ __forceinline__ int rec_type(double w) { return (int)(__double_as_longlong(w) & 0xff); }
__device__ __forceinline__ int rec_mask(double w) { return (int)((__double_as_longlong(w) >> 8) & 0xffffffffLL); }
__host__ __device__ __forceinline__ long long pack_bits(int type, int mask) { return ((long long)(unsigned)mask << 8) | (long long)(type & 0xff); }

// sphere/wall contact with the wall's own model selection (uniform run-time switch)
__device__ __noinline__ void wall_chain(const StepP &P, const ModelP &M, const Contact &c, double (&h)[3], double (&g)[3], bool su, ContactOut &o)
{
  const int key = M.normal * 4 + M.rolling;
  switch (key) {
    case N_HERTZ * 4 + R_OFF: contact_chain<N_HERTZ, R_OFF, true>(P, M, c, h, g, su, o); break;
    case N_HERTZ * 4 + R_CDT: contact_chain<N_HERTZ, R_CDT, true>(P, M, c, h, g, su, o); break;
    case N_HERTZ * 4 + R_EPSD: contact_chain<N_HERTZ, R_EPSD, true>(P, M, c, h, g, su, o); break;
    case N_HERTZ * 4 + R_EPSD2: contact_chain<N_HERTZ, R_EPSD2, true>(P, M, c, h, g, su, o); break;
    case N_HOOKE * 4 + R_OFF: contact_chain<N_HOOKE, R_OFF, true>(P, M, c, h, g, su, o); break;
    case N_HOOKE * 4 + R_CDT: contact_chain<N_HOOKE, R_CDT, true>(P, M, c, h, g, su, o); break;
    case N_HOOKE * 4 + R_EPSD: contact_chain<N_HOOKE, R_EPSD, true>(P, M, c, h, g, su, o); break;
    default: contact_chain<N_HOOKE, R_EPSD2, true>(P, M, c, h, g, su, o); break;
  }
}

// primitive walls of one particle: fix_wall_gran.cpp:988-1121, fix_wall_gran_base.h:159-367,
// primitive_wall_definitions.h:128-203
// Pre-pass kernel over the compact list of particles that are candidates of at least one primitive
// wall (a thin layer of the bed): keeps ~200 lines of rarely needed code and its registers out of
// the hot kernel.  Writes the wall force/torque of candidate c to fw[0..5][c]; k_step adds it.
__global__ void __launch_bounds__(128) k_walls(const StepP P)
{
  const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (cidx >= P.nwc || step_gated(P)) return;
  const int i = P.wlist[cidx];
  const double4 xi = ldg4(P.xr + i), vi = ldg4(P.vm + i), wi = ldg4(P.wt + i);
  const int itype = rec_type(wi.w);
  const bool su = (P.mode != MODE_SETUP);
  double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
  const double4 xh = P.xh[i];
  const unsigned wbits = (unsigned)(__double_as_longlong(xh.w) & 0xffffffffLL);
  const unsigned cand = wbits & 0xffffu;
  unsigned valid = wbits >> 16;
  const unsigned valid0 = valid;
  const double pos[3] = {xi.x, xi.y, xi.z};
  const double r = xi.w;
  for (int w = 0; w < P.nwalls; w++) {
    if (!((cand >> w) & 1u)) continue;
    const WallP &W = P.walls[w];
    double delta[3] = {0., 0., 0.}, deltan;
    if (W.wtype < 3) {
      const int d = W.wtype;
      const double p = W.param[0];
      delta[d] = p - pos[d];
      deltan = pos[d] > p ? pos[d] - p - r : p - pos[d] - r;
    } else {
      const int dd = W.wtype - 3, iy = (dd + 1) % 3, iz = (dd + 2) % 3;
      const double R = W.param[0];
      const double dy = pos[iy] - W.param[1], dz = pos[iz] - W.param[2];
      const double dist = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz)));
      if (dist == 0.0) deltan = 0.0;
      else if (dist > R) { deltan = dist - R - r; const double fact = (dist - R) / dist; delta[iy] = -dy * fact; delta[iz] = -dz * fact; }
      else { deltan = R - dist - r; const double fact = (R - dist) / dist; delta[iy] = dy * fact; delta[iz] = dz * fact; }
    }
    if (deltan > P.cutneighmax) continue;
    const int wd = W.m.dnum;
    if (deltan <= 0 || deltan < (P.cdf - 1.0) * r) {
      if (deltan <= 0) {
        Contact c;
        c.dx = -delta[0]; c.dy = -delta[1]; c.dz = -delta[2];
        c.radi = r; c.radj = 0.0; c.radsum = r; c.deltan_in = -deltan;
        c.r = c.radi - c.deltan_in; c.rinv = 1.0 / c.r;
        c.meff = vi.w; c.mi = vi.w; c.mj = 0.0;
        c.vi[0] = vi.x; c.vi[1] = vi.y; c.vi[2] = vi.z;
        c.wi[0] = wi.x; c.wi[1] = wi.y; c.wi[2] = wi.z;
        c.vj[0] = c.vj[1] = c.vj[2] = 0.0; c.wj[0] = c.wj[1] = c.wj[2] = 0.0;
        if (W.shear) {
          if (W.shearAxis >= 0) {
            const int dd = W.wtype - 3;
            double rd[3] = {0., 0., 0.};
            rd[(dd + 1) % 3] = pos[(dd + 1) % 3] - W.param[1];
            rd[(dd + 2) % 3] = pos[(dd + 2) % 3] - W.param[2];
            c.vj[0] = W.axisVec[1] * rd[2] - W.axisVec[2] * rd[1];
            c.vj[1] = W.axisVec[2] * rd[0] - W.axisVec[0] * rd[2];
            c.vj[2] = W.axisVec[0] * rd[1] - W.axisVec[1] * rd[0];
          } else c.vj[W.shearDim] = W.vshear;
        }
        c.itype = itype; c.jtype = W.atom_type;
        double h[3] = {0., 0., 0.}, g[3] = {0., 0., 0.};
        if ((valid >> w) & 1u) {
#pragma unroll
          for (int d = 0; d < 3; d++) {
            if (W.m.tangential) h[d] = P.whist[(size_t)(W.hist_row + W.m.off_shear + d) * P.cap + i];
            if (W.m.off_roll >= 0) g[d] = P.whist[(size_t)(W.hist_row + W.m.off_roll + d) * P.cap + i];
          }
        }
        ContactOut o;
        wall_chain(P, W.m, c, h, g, su, o);
        F[0] += o.F[0]; F[1] += o.F[1]; F[2] += o.F[2];
        T[0] += o.Ti[0]; T[1] += o.Ti[1]; T[2] += o.Ti[2];
        if (su && wd) {
#pragma unroll
          for (int d = 0; d < 3; d++) {
            if (W.m.tangential) P.whist[(size_t)(W.hist_row + W.m.off_shear + d) * P.cap + i] = h[d];
            if (W.m.off_roll >= 0) P.whist[(size_t)(W.hist_row + W.m.off_roll + d) * P.cap + i] = g[d];
          }
          valid |= (1u << w);
        }
      } else valid &= ~(1u << w);  // surfacesClose: history zeroed (tangential_model_history.h:428-440)
    } else valid &= ~(1u << w);  // candidate but apart: history zeroed (fix_wall_gran.cpp:1117-1119)
  }
  if (valid != valid0) {
    const long long nb = (__double_as_longlong(xh.w) & ~0xffffffffLL) | (long long)(cand | (valid << 16));
    P.xh[i].w = __longlong_as_double(nb);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) { P.fw[(size_t)d * P.nwcap + cidx] = F[d]; P.fw[(size_t)(3 + d) * P.nwcap + cidx] = T[d]; }
}

// ---- tuning constants of k_step (each set by measurement on the 4,194,304-sphere bed; profiles/)
#ifndef DEM_CMAX
#define DEM_CMAX 12  // touching entries per particle staged in shared memory (more: evaluated by the owner on the spot)
#endif
#ifndef DEM_OWNR
#define DEM_OWNR 12  // upper limit of contacts per particle evaluated by the particle's own lane before the cooperative deal;
#endif               // 0 = cooperative deal only (1.252 ms), 12 with the cost rule below 1.125 ms, fixed 5: 1.118 ms, all own: 1.33 ms
#ifndef DEM_COST_OWN
#define DEM_COST_OWN 3   // relative cost of an own-lane round ...
#define DEM_COST_COOP 5  // ... and of a cooperative round (r01t / r01u)
#endif
#ifndef DEM_SWEEPW
#define DEM_SWEEPW 5  // row entries per sweep pass = position gathers in flight per thread (4: 1.125 ms, 5: 1.082 ms, 6: 1.215 ms, 8: 1.31 ms -- spills)
#endif
#ifndef DEM_CPREFETCH
#define DEM_CPREFETCH 1  // L2 prefetch of a staged contact's operands: 0 none (0.900 ms), 1 history rows (0.875), 2 history + partner v|m, omega|bits
                         // (0.906: since v8 the load/store pipe is the binding unit and every prefetch is one more wavefront per lane)
#endif
#ifndef DEM_PIPE
#define DEM_PIPE 0   // contact rounds software pipelined (operands of the next round in flight during the evaluation)
#endif
#ifndef DEM_STEP_MINBLOCKS
#define DEM_STEP_MINBLOCKS 5    // 96 registers; 4 (128 registers) 1.45 ms, 6 (80 registers, spills) 1.42 ms
#endif
#ifndef DEM_STEP_WAVE_PREFETCH
#define DEM_STEP_WAVE_PREFETCH 100  // blocks ahead whose streaming inputs are pulled towards L2 (v8: 50 0.870, 100 0.870, 150 0.873, 200 0.875, 400 0.887 ms; 0 = off: 1.31 ms on v7)
#endif

// one list entry w of particle i (block-local index q, operands from the block's shared-memory records): evaluated in MY
// orientation (see pair_chain) if the spheres touch; history records are stored in the canonical orientation "lower tag
// first" (sign bit flipped on load/store when the partner is the first body).  s_nh[q] = the particle's count of history slots
// in use (a shared-memory counter: in the cooperative phase another lane may be serving the particle).
// STD (compile-time): the deck uses the reference's default sub-model settings (tangential history with damping, no
// limitForce, no torsion torque, nktv2p == 1, contact_distance_factor 1, no per-contact output) -- the uniform run-time
// switches and the generic history-row arithmetic drop out of the hot loop.
__device__ __forceinline__ double flip_if(double v, unsigned flip)
{  // v with its sign bit XORed by `flip` (0 or 0x80000000): history sign without a multiplication
  return __hiloint2double(__double2hiint(v) ^ (int)flip, __double2loint(v));
}
__device__ __forceinline__ double4 rec_get(const double2 (*s_rec)[128], int a, int q)
{
  const double2 lo = s_rec[2 * a][q], hi = s_rec[2 * a + 1][q];
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void rec_put(double2 (*s_rec)[128], int a, int q, const double4 &v)
{
  s_rec[2 * a][q] = make_double2(v.x, v.y); s_rec[2 * a + 1][q] = make_double2(v.z, v.w);
}
// operands of one list entry: partner records + the pair's history records (loaded a round ahead, see k_step)
struct ItemOps { double4 xj, vj, wj, hs, hr; };
__device__ __forceinline__ double4 ld4(const double4 *p)
{  // plain (coherent) 256-bit load: history records are rewritten by this kernel
  double4 v;
#if DEM_STREAM_LD
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#else
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#endif
  return v;
}
template <int ROLLING, bool STD>
__device__ __forceinline__ void item_load(const StepP &P, int i, unsigned w, bool valid, ItemOps &o)
{
  constexpr bool HAS_ROLL_HIST = (ROLLING == R_EPSD || ROLLING == R_EPSD2);
  const int hrec = STD ? (HAS_ROLL_HIST ? 2 : 1) : P.pm.hrec;
  const int rec_shear = STD ? 0 : P.pm.rec_shear, rec_roll = STD ? 1 : P.pm.rec_roll;
  const bool tangential = STD ? true : (P.pm.tangential != 0);
  o.hs = make_double4(0., 0., 0., 0.); o.hr = make_double4(0., 0., 0., 0.);
  if (!valid) { o.xj = o.vj = o.wj = o.hs; return; }
  const int j = (int)(w & NBR_IDX);
  const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
  o.xj = ldg4(P.xr + j); o.vj = ldg4(P.vm + j); o.wj = ldg4(P.wt + j);
  if (slot >= 0) {
    const double4 *hp = P.hist + (size_t)(slot * hrec) * P.lcap + i;
    if (tangential) o.hs = ld4(hp + (size_t)rec_shear * P.lcap);
    if (HAS_ROLL_HIST) o.hr = ld4(hp + (size_t)rec_roll * P.lcap);
  }
}
template <int NORMAL, int ROLLING, bool ONE, bool F32, bool STD>
__device__ __forceinline__ void pair_item(const StepP &P, int i, int q, unsigned w, const ItemOps &o, const double2 (*s_rec)[128], int *s_nh, bool su,
                                          double (&Fc)[3], double (&Tc)[3])
{
  constexpr bool HAS_ROLL_HIST = (ROLLING == R_EPSD || ROLLING == R_EPSD2);
  const int hrec = STD ? (HAS_ROLL_HIST ? 2 : 1) : P.pm.hrec;
  const int rec_shear = STD ? 0 : P.pm.rec_shear, rec_roll = STD ? 1 : P.pm.rec_roll;
  const bool tangential = STD ? true : (P.pm.tangential != 0);
  const int j = (int)(w & NBR_IDX);
  int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
  const bool had = slot >= 0;
  const double4 &xj = o.xj, &vj = o.vj, &wj = o.wj, &hs = o.hs, &hr = o.hr;
  const double4 xi = rec_get(s_rec, 0, q), vi = rec_get(s_rec, 1, q), wi = rec_get(s_rec, 2, q);
  const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
  const double rsq = sq3_rn(dx, dy, dz);
  const double radsum = xi.w + xj.w;
  if (!(rsq < __dmul_rn(radsum, radsum))) return;  // pair_gran_base.h:358 (entries that were staged by the sweep always pass)
  const unsigned flip = w & NBR_JFIRST;
  double h[3] = {flip_if(hs.x, flip), flip_if(hs.y, flip), flip_if(hs.z, flip)}, g[3] = {flip_if(hr.x, flip), flip_if(hr.y, flip), flip_if(hr.z, flip)};
  if (F32)
    pair_chain_f32<NORMAL, ROLLING, ONE>(P, P.pm, xi, vi, wi, xj, vj, wj, rec_type(wi.w), rec_type(wj.w), rec_mask(wi.w), rec_mask(wj.w), dx, dy, dz, rsq, h, g, su, Fc, Tc, nullptr);
  else
    pair_chain<NORMAL, ROLLING, ONE, STD>(P, P.pm, xi, vi, wi, xj, vj, wj, rec_type(wi.w), rec_type(wj.w), rec_mask(wi.w), rec_mask(wj.w), dx, dy, dz, rsq, h, g, su, Fc, Tc);
  if (!had) {  // first touch since the last rebuild: the contact flag becomes != 0 and stays
    const int s = atomicAdd(s_nh + q, 1);
    if (s < P.hslots) {
      slot = s;
      const int nn = P.numneigh[i] & 0xffff;
      for (int k = 0; k < nn; k++)  // rare path: find the row entry of this partner and tag it with its slot
        if ((P.nbr[(size_t)k * P.lcap + i] & NBR_IDX) == (unsigned)j) { P.nbr[(size_t)k * P.lcap + i] = w | ((unsigned)(slot + 1) << NBR_SLOT_SHIFT); break; }
    } else { atomicSub(s_nh + q, 1); ((volatile int *)P.flag)[1] = 1; }
  }
  if (hrec && slot >= 0 && (su || !had)) {
    double4 *hp = P.hist + (size_t)(slot * hrec) * P.lcap + i;
    if (tangential) st4(hp + (size_t)rec_shear * P.lcap, make_double4(flip_if(h[0], flip), flip_if(h[1], flip), flip_if(h[2], flip), 0.));
    if (HAS_ROLL_HIST) st4(hp + (size_t)rec_roll * P.lcap, make_double4(flip_if(g[0], flip), flip_if(g[1], flip), flip_if(g[2], flip), 0.));
  }
  if (!STD && P.cout && slot >= 0) {  // option contact_output, setup / last step only (null otherwise): see k_contact_fill
    double4 *cp = P.cout + ((size_t)slot * P.lcap + i) * 2;
    st4(cp, make_double4(Fc[0], Fc[1], Fc[2], Tc[0]));
    st4(cp + 1, make_double4(Tc[1], Tc[2], P.serial, 0.));
  }
}

// start the memory accesses a staged contact will need, without holding registers
template <bool STD, bool HAS_ROLL_HIST>
__device__ __forceinline__ void prefetch_contact(const StepP &P, int i, unsigned w)
{
  const int j = (int)(w & NBR_IDX);
  if (DEM_CPREFETCH >= 2) { prefetch_l2(P.vm + j); prefetch_l2(P.wt + j); }
  const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
  if (DEM_CPREFETCH >= 1 && slot >= 0) {
    const int hrec = STD ? (HAS_ROLL_HIST ? 2 : 1) : P.pm.hrec;
    const double4 *hp = P.hist + (size_t)(slot * hrec) * P.lcap + i;
    for (int r = 0; r < hrec; r++) prefetch_l2(hp + (size_t)r * P.lcap);
  }
}

// owner epilogue of a step: gravity, primitive/mesh wall force of the pre-pass, freeze, (optional) force output,
// final_integrate(n) + initial_integrate(n+1), rebuild trigger.  fix_gravity.cpp:331-339, fix_freeze.cpp:132-144,
// fix_nve_sphere.cpp:134-244, neighbor.cpp:1425-1466
// XF: the kernel also applies the fix addforce / fix viscous list (never compiled into the specialised hot kernel)
template <bool XF = false>
__device__ __forceinline__ bool step_epilogue(const StepP &P, int i, const double4 &xi, const double4 &vi, const double4 &wi, double *F, double *T)
{
  bool trig = false;
  const int imask = rec_mask(wi.w);
  if (P.have_g && (imask & 1)) { F[0] += vi.w * P.g[0]; F[1] += vi.w * P.g[1]; F[2] += vi.w * P.g[2]; }
  if (P.nwc) {  // wall contacts were evaluated by the k_walls / k_mesh_step pre-passes
    const unsigned widx = (unsigned)(__double_as_longlong(P.xh[i].w) >> 32);
    if (widx) {
#pragma unroll
      for (int d = 0; d < 3; d++) { F[d] += P.fw[(size_t)d * P.nwcap + widx - 1]; T[d] += P.fw[(size_t)(3 + d) * P.nwcap + widx - 1]; }
    }
  }
  if (XF) {
    for (int q = 0; q < P.nxf; q++) {
      const XForce &X = P.xf[q];
      if (!(imask & X.bit)) continue;
      if (X.kind == 0) { F[0] += X.v[0]; F[1] += X.v[1]; F[2] += X.v[2]; }
      else { F[0] -= X.v[0] * vi.x; F[1] -= X.v[0] * vi.y; F[2] -= X.v[0] * vi.z; }
    }
  }
  if (imask & P.freezebit) { F[0] = F[1] = F[2] = 0.0; T[0] = T[1] = T[2] = 0.0; }
  if (P.mode != MODE_STEP) {  // forces are only materialised when somebody will read them
    P.f[i] = F[0]; P.f[P.cap + i] = F[1]; P.f[2 * (size_t)P.cap + i] = F[2];
    P.tq[i] = T[0]; P.tq[P.cap + i] = T[1]; P.tq[2 * (size_t)P.cap + i] = T[2];
  }
  if (P.mode != MODE_SETUP) {
    double4 xo = xi, vo = vi, wo = wi;
    if (imask & P.integbit) {
      const double dtfm = P.dtf / vi.w;
      const double dtir = P.dtfrot / (xi.w * xi.w * vi.w);
      // final_integrate of this step
      vo.x += dtfm * F[0]; vo.y += dtfm * F[1]; vo.z += dtfm * F[2];
      wo.x += dtir * T[0]; wo.y += dtir * T[1]; wo.z += dtir * T[2];
      if (P.mode == MODE_STEP) {  // initial_integrate of the next step
        vo.x += dtfm * F[0]; vo.y += dtfm * F[1]; vo.z += dtfm * F[2];
        xo.x += P.dtv * vo.x; xo.y += P.dtv * vo.y; xo.z += P.dtv * vo.z;
        wo.x += dtir * T[0]; wo.y += dtir * T[1]; wo.z += dtir * T[2];
      }
    }
    st4(P.xr_o + i, xo); st4(P.vm_o + i, vo); st4(P.wt_o + i, wo);
    if (P.img_first) {  // fused ghost push: every copy of this particle (ghost slots of the neighbour ranks, periodic images here)
      int k = P.img_first[i];
      if (k >= 0) {
        const ImgP &I = *P.img;
        for (; k >= 0;) {
          const int4 e = P.img_tab[k];
          k = e.z;
          const int t = (e.x >> 28) & 3, slot = e.x & 0x0fffffff, b = I.pcur0[t] ^ P.img_par;
          double4 xs = xo;
          if (e.y & 3) xs.x += (e.y & 1) ? I.prd[0] : -I.prd[0];
          if (e.y & 12) xs.y += (e.y & 4) ? I.prd[1] : -I.prd[1];
          if (e.y & 48) xs.z += (e.y & 16) ? I.prd[2] : -I.prd[2];
          st4(I.bx[t][b] + slot, xs); st4(I.bv[t][b] + slot, vo); st4(I.bw[t][b] + slot, wo);
        }
      }
    }
    if (P.mode == MODE_STEP) {
      const double4 xh = P.xh[i];
      trig = sq3_rn(xo.x - xh.x, xo.y - xh.y, xo.z - xh.z) > P.trigsq;
    }
  }
  return trig;
}

// ----------------------------------------------------------------------------------------
// THE hot kernel: one launch == one timestep of the owned particles of this GPU.
//   verlet.cpp:264-391 (order of operations), pair_gran_base.h:257-496 (pair loop),
//   fix_gravity.cpp:331-339, fix_freeze.cpp:132-144, fix_nve_sphere.cpp:134-244,
//   neighbor.cpp:1425-1466 (rebuild trigger)
// Phases per warp (32 consecutive particles):
//  (1) every lane walks its own row in passes of DEM_SWEEPW entries: the neighbour words, as many independent position gathers in flight,
//      the touch verdicts; the words of the touching entries go (from registers) into the lane's column of a shared-memory
//      staging area, and the fetches of what their evaluation will need are started towards L2;
//  (2) ONE loop of contact rounds whose body (pair_item) exists once in the code (the kernel is instruction-fetch sensitive:
//      with three inlined copies 13 % of the stall samples were "no instruction"):
//      (a) own rounds: every lane evaluates the first contacts of its own particle -- no owner search, no result round trip,
//          nearly all lanes busy; how many such rounds is decided per warp by a cost rule;
//      (b) cooperative rounds: the uneven remainder -- the (particle, contact) items of the 32 particles are dealt out
//          round-robin to the 32 lanes, each item's force/torque goes through a shared-memory slot back to the owning lane,
//          which adds its items in list order;
//      (c) (rare) rounds for the touching entries beyond the DEM_CMAX staged ones, re-read from the row by their owner.
//      A particle's sum runs in list order and does not depend on which lane evaluated what: runs are bit reproducible;
//  (3) the owner adds gravity / wall force, integrates, writes its records and votes on the rebuild.
// Alternatives that were built, found parity-green and measured slower on the 4.19M bed (git history):
// partner records from shared memory, one evaluation per in-warp pair, a separate sweep kernel, a software-pipelined
// contact phase, an owner list in two launches / as a persistent wavefront.
template <int NORMAL, int ROLLING, bool ONE, bool F32 = false, bool STD = false>
__global__ void __launch_bounds__(128, DEM_STEP_MINBLOCKS) k_step(const StepP P)
{  // F32: option fp32 -- the contact law in single precision (pair_chain_f32); state, geometry, sums and integration stay fp64
  constexpr bool HAS_ROLL_HIST = (ROLLING == R_EPSD || ROLLING == R_EPSD2);
  __shared__ unsigned s_w[DEM_CMAX][128];
  __shared__ double2 s_rec[6][128];   // own records of the block's particles (x|r, v|m, omega|bits) as 16-byte halves: LDS.128 at a 16-byte lane stride is conflict free
  __shared__ double s_res[6][4 * 32];  // per warp: force / torque of the items of the current cooperative round
  __shared__ int s_off[128], s_nh[128];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wb = tid & ~31;
  if (step_gated(P)) return;
  const bool active = i < P.nlocal;
  const bool su = (P.mode != MODE_SETUP);
  bool trig = false;
  double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
  int nc = 0, nh0 = 0, nn = 0, kov = 0x7fffffff;
#if DEM_STEP_WAVE_PREFETCH > 0
  {  // pull the streaming inputs of the block that will run one wave later towards L2
    const int ip = i + DEM_STEP_WAVE_PREFETCH * 128;
    if (ip < P.nlocal) {
      prefetch_l2(P.xr + ip); prefetch_l2(P.vm + ip); prefetch_l2(P.wt + ip); prefetch_l2(P.xh + ip);
      if (lane < 12) prefetch_l2(P.nbr + (size_t)lane * P.lcap + (ip - lane));
      else if (lane == 12) prefetch_l2(P.numneigh + (ip - lane));
      else if (lane == 13 && P.img_first) prefetch_l2(P.img_first + (ip - lane));
    }
  }
#endif
  {
    double4 xi = make_double4(0., 0., 0., 0.), vi = xi, wi = xi;
    if (active) { xi = ldg4s(P.xr + i); vi = ldg4s(P.vm + i); wi = ldg4s(P.wt + i); }
    rec_put(s_rec, 0, tid, xi); rec_put(s_rec, 1, tid, vi); rec_put(s_rec, 2, tid, wi);
    if (active && P.have_pair) {
      const int nnw = P.numneigh[i];
      nn = nnw & 0xffff;
      nh0 = (nnw >> 16) & 0xffff;
    }
    s_nh[tid] = nh0;
    for (int k0 = 0; k0 < ((!STD && (P.debug & 2)) ? 0 : nn); k0 += DEM_SWEEPW) {
      // (1) one pass = DEM_SWEEPW row entries: words, gathers, verdicts -- branch free, then the staging of the touching ones
      unsigned wv[DEM_SWEEPW];
      double4 xv[DEM_SWEEPW];
#pragma unroll
      for (int u = 0; u < DEM_SWEEPW; u++) wv[u] = (k0 + u < nn) ? (DEM_STREAM_LD ? __ldcg(P.nbr + (size_t)(k0 + u) * P.lcap + i) : P.nbr[(size_t)(k0 + u) * P.lcap + i]) : (unsigned)i;  // past the row's end: myself (never touches)
#pragma unroll
      for (int u = 0; u < DEM_SWEEPW; u++) xv[u] = ldg4(P.xr + (wv[u] & NBR_IDX));
      unsigned touch = 0u, close = 0u;
#pragma unroll
      for (int u = 0; u < DEM_SWEEPW; u++) {
        const double rsq = sq3_rn(xi.x - xv[u].x, xi.y - xv[u].y, xi.z - xv[u].z);
        const double radsum = xi.w + xv[u].w;
        bool t = (k0 + u < nn) && rsq < __dmul_rn(radsum, radsum);
        touch |= (unsigned)t << u;
        if (!STD && P.cdf > 1.0) close |= (unsigned)((k0 + u < nn) && !t && (wv[u] & NBR_HIST) && rsq < P.cdfsq * radsum * radsum) << u;
      }
#pragma unroll
      for (int u = 0; u < DEM_SWEEPW; u++) {
        if (!((touch >> u) & 1u)) continue;
        if (nc < DEM_CMAX) { s_w[nc++][tid] = wv[u]; prefetch_contact<STD, HAS_ROLL_HIST>(P, i, wv[u]); }
        else kov = min(kov, k0 + u);  // more than DEM_CMAX contacts (rare): rounds (c) re-read the row from here
      }
      while (close) {  // surfacesClose: tangential/rolling history zeroed, flag stays, pair_gran_base.h:420-423
        const int u = __ffs((int)close) - 1;
        close &= close - 1;
        const unsigned w = P.nbr[(size_t)(k0 + u) * P.lcap + i];
        const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
        // (the hertz / hooke normal models leave their bit of the pair's contact flags set in surfacesClose, normal_model_hertz.h:410-413:
        //  the pair keeps its row and stays flagged)
        for (int r = 0; r < P.pm.hrec; r++) st4(P.hist + (size_t)(slot * P.pm.hrec + r) * P.lcap + i, make_double4(0., 0., 0., 0.));
      }
    }
    if (!STD && (P.debug & 1)) { nc = 0; kov = 0x7fffffff; }
  }
  // number of own rounds: minimises (own rounds) x DEM_COST_OWN + (cooperative rounds of 32 items that remain) x DEM_COST_COOP
  // for this warp -- an own round costs the same whatever the number of busy lanes, a cooperative round (owner search,
  // result round trip, owner accumulation) costs more but is always full.
  int ownr = 0;
  {
    int best = 0x7fffffff;
#pragma unroll 1
    for (int r = 0; r <= DEM_OWNR; r++) {
      const int rem = __reduce_add_sync(0xffffffffu, max(nc - r, 0));
      const int cost = r * DEM_COST_OWN + ((rem + 31) >> 5) * DEM_COST_COOP;
      if (cost < best) { best = cost; ownr = r; }
      if (rem == 0) break;
    }
  }
  // cooperative deal of the remaining items: item t belongs to the last lane whose first item is <= t
  const int ncc = max(nc - ownr, 0);
  int incl = ncc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const int excl = incl - ncc;
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  s_off[tid] = excl;
  __syncwarp();
  const int rcoop = ownr + ((total + 31) >> 5);
  // (2) the contact rounds, software pipelined: the operands of round r+1 (partner records, history) are loaded while round
  //     r is evaluated -- a lane's own records come from shared memory when they are used, which is what leaves the registers
  //     for the second operand set
  auto select = [&](int rnd, int &q, unsigned &w) -> bool {
    bool valid;
    q = tid; w = 0u;
    if (rnd < ownr) { valid = rnd < nc; if (valid) w = s_w[rnd][tid]; }
    else if (rnd < rcoop) {
      const int t = ((rnd - ownr) << 5) + lane;
      valid = t < total;
      if (valid) {
        int p = 0;
#pragma unroll
        for (int s = 16; s; s >>= 1) if (s_off[wb + p + s] <= t) p += s;
        q = wb + p;
        w = s_w[ownr + t - s_off[q]][q];
      }
    } else { valid = kov < nn; if (valid) { w = P.nbr[(size_t)kov * P.lcap + i]; kov++; } }
    return valid;
  };
  int q; unsigned w;
  bool valid = select(0, q, w);
  ItemOps oc;
  item_load<ROLLING, STD>(P, i - tid + q, w, valid, oc);
#pragma unroll 1
  for (int rnd = 0;; rnd++) {
    if (rnd >= rcoop && !__any_sync(0xffffffffu, valid)) break;
#if DEM_PIPE
    int qn; unsigned wn;
    const bool validn = select(rnd + 1, qn, wn);
    ItemOps on;
    item_load<ROLLING, STD>(P, i - tid + qn, wn, validn, on);
#endif
    double rF[3] = {0., 0., 0.}, rT[3] = {0., 0., 0.};
    if (valid) pair_item<NORMAL, ROLLING, ONE, F32, STD>(P, i - tid + q, q, w, oc, s_rec, s_nh, su, rF, rT);
    if (rnd >= ownr && rnd < rcoop) {  // cooperative round: results are parked in shared memory and each owner adds its items (in list order)
      const int t0 = (rnd - ownr) << 5;
      const int sl = wb + lane;
#pragma unroll
      for (int d = 0; d < 3; d++) { s_res[d][sl] = rF[d]; s_res[3 + d][sl] = rT[d]; }
      __syncwarp();
      const int qe = min(incl, t0 + 32);
      for (int k = max(excl, t0); k < qe; k++) {
        const int so = wb + (k - t0);
#pragma unroll
        for (int d = 0; d < 3; d++) { F[d] += s_res[d][so]; T[d] += s_res[3 + d][so]; }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int d = 0; d < 3; d++) { F[d] += rF[d]; T[d] += rT[d]; }
    }
#if DEM_PIPE
    q = qn; w = wn; valid = validn; oc = on;
#else
    valid = select(rnd + 1, q, w);
    item_load<ROLLING, STD>(P, i - tid + q, w, valid, oc);
#endif
  }
  // (3) owner epilogue
  if (active) {
    if (P.have_pair) { const int nh = s_nh[tid]; if (nh != nh0) P.numneigh[i] = nn | (nh << 16); }
    trig = step_epilogue<!STD>(P, i, rec_get(s_rec, 0, tid), rec_get(s_rec, 1, tid), rec_get(s_rec, 2, tid), F, T);
  }
  if (__any_sync(0xffffffffu, trig) && (threadIdx.x & 31) == 0) *((volatile int *)P.flag) = 1;
}


// ----------------------------------------------------------------------------------------
// Step kernel of the bonded-sphere decks (pair_style gran ... cohesion bond | bond/nonlinear).  Same contract as k_step;
// differences: (a) the contact-distance factor is > 1, so entries inside the band (radsum*cdf) are evaluated although the
// spheres do not touch: bonded pairs pull across the gap, un-bonded ones are bond candidates on the creation step;
// (b) every pair is evaluated in the canonical orientation (lower tag = first body) by both owners -- see bond_eval;
// (c) a history row is nbrec + 1 (+1) records; the double after the bond values is the row's sticky flag S: the reference's
// contact_flags stay != 0 after the first touch (CONTACT_NORMAL_MODEL is never cleared) and are reset to 1 for rows kept by
// a rebuild (neigh_gran.cpp:596-612), while CONTACT_COHESION_MODEL follows the bond; "flag != 0" == (S != 0 || bondFlag != 0).
// Chain order per contact_models.h:228-253: surface, normal, cohesion, tangential, rolling.
#ifndef DEM_BOND_MINBLOCKS
#define DEM_BOND_MINBLOCKS 2
#endif
template <int NORMAL, int ROLLING, int COH>
__global__ void __launch_bounds__(128, DEM_BOND_MINBLOCKS) k_step_bond(const StepP P)
{
  constexpr bool HAS_ROLL_HIST = (ROLLING == R_EPSD || ROLLING == R_EPSD2);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool trig = false;
  if (step_gated(P)) return;
  if (i < P.nlocal) {
    const ModelP &M = P.pm;
    const double4 xi = ldg4(P.xr + i), vi = ldg4(P.vm + i), wi = ldg4(P.wt + i);
    const int itype = rec_type(wi.w), imask = rec_mask(wi.w);
    const bool su = (P.mode != MODE_SETUP);
    const bool create_step = su && (M.createAlways || P.ntimestep == P.tsCreateBond);
    double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
    const int nnw = P.numneigh[i];
    const int nn = nnw & 0xffff;
    int nh = (nnw >> 16) & 0xffff;
    const int nh0 = nh;
    constexpr int NB = (COH == C_BOND ? 14 : 28);  // == M.nbond (compile-time so that the history row lives in registers)
    for (int k0 = 0; k0 < nn; k0 += 32) {
    // (a) branch-free sweep of up to 32 row entries, 8 position gathers in flight: which entries need the contact chain
    //     (touching, or inside the contact-distance band and bonded / bond candidates on the creation step)
    unsigned need = 0u;
    const int kn = min(32, nn - k0);
#pragma unroll 8
    for (int kk = 0; kk < kn; kk++) {
      const unsigned w = P.nbr[(size_t)(k0 + kk) * P.lcap + i];
      const double4 xj = ldg4(P.xr + (w & NBR_IDX));
      const double rsq = sq3_rn(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
      const double radsum = xi.w + xj.w;
      const bool touch = rsq < __dmul_rn(radsum, radsum);
      const bool inband = touch || rsq < P.cdfsq * radsum * radsum;
      need |= (unsigned)(inband && (touch || (w & NBR_HIST) || create_step)) << kk;
    }
    // (b) the selected entries
    while (need) {
      const int k = k0 + __ffs((int)need) - 1;
      need &= need - 1;
      unsigned w = P.nbr[(size_t)k * P.lcap + i];
      const int j = (int)(w & NBR_IDX);
      const double4 xj = ldg4(P.xr + j);
      const double dxm = xi.x - xj.x, dym = xi.y - xj.y, dzm = xi.z - xj.z;  // me - partner
      const double rsq = sq3_rn(dxm, dym, dzm);
      const double radsum = xi.w + xj.w;
      const bool touch = rsq < __dmul_rn(radsum, radsum);
      int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
      const double4 vj = ldg4(P.vm + j), wj = ldg4(P.wt + j);
      const bool jfirst = (w & NBR_JFIRST) != 0;
      // canonical operands: a = first body (lower tag), b = second
      const double4 &xa = jfirst ? xj : xi, &xb = jfirst ? xi : xj, &va = jfirst ? vj : vi, &vb = jfirst ? vi : vj, &wa = jfirst ? wj : wi, &wb = jfirst ? wi : wj;
      const double delta[3] = {xa.x - xb.x, xa.y - xb.y, xa.z - xb.z};
      const int ta = rec_type(wa.w), tb = rec_type(wb.w);
      // history row -> registers
      double H[28], S = 0.0, h[3] = {0., 0., 0.}, g[3] = {0., 0., 0.};
#pragma unroll
      for (int d = 0; d < 28; d++) H[d] = 0.0;
      const bool had = slot >= 0;
      if (had) {
        const double4 *hp = P.hist + (size_t)(slot * M.hrec) * P.lcap + i;
#pragma unroll
        for (int r = 0; r < (COH == C_BOND ? 4 : 8); r++) {
          const double4 v = hp[(size_t)(M.rec_bond + r) * P.lcap];
          if (4 * r < NB) H[4 * r] = v.x; else if (4 * r == NB) S = v.x;
          if (4 * r + 1 < NB) H[4 * r + 1] = v.y; else if (4 * r + 1 == NB) S = v.y;
          if (4 * r + 2 < NB) H[4 * r + 2] = v.z; else if (4 * r + 2 == NB) S = v.z;
          if (4 * r + 3 < NB) H[4 * r + 3] = v.w; else if (4 * r + 3 == NB) S = v.w;
        }
        if (M.tangential) { const double4 v = hp[(size_t)M.rec_shear * P.lcap]; h[0] = v.x; h[1] = v.y; h[2] = v.z; }
        if (HAS_ROLL_HIST) { const double4 v = hp[(size_t)M.rec_roll * P.lcap]; g[0] = v.x; g[1] = v.y; g[2] = v.z; }
      }
      const double S0 = S;
      // cohesion model first (it only needs kinematics + its own history); force on the first body
      double Fb[3] = {0., 0., 0.}, Tbi[3] = {0., 0., 0.}, Tbj[3] = {0., 0., 0.};
      const double xav[3] = {xa.x, xa.y, xa.z}, vav[3] = {va.x, va.y, va.z}, vbv[3] = {vb.x, vb.y, vb.z}, wav[3] = {wa.x, wa.y, wa.z}, wbv[3] = {wb.x, wb.y, wb.z};
      int bev = 0;
      const bool bonded = bond_eval<COH>(P, M, delta, rsq, xa.w, xb.w, xav, vav, vbv, wav, wbv, ta, tb, su, H, Fb, Tbi, Tbj, bev);
      if (bev && P.bondc && !jfirst) {  // compute bond/counter: the lower-tag particle of a pair counts its events
        if (bev & 1) atomicAdd(P.bondc, 1ULL);
        if (bev & 2) atomicAdd(P.bondc + 1, 1ULL);
      }
      double Fa[3] = {0., 0., 0.}, Ta[3] = {0., 0., 0.}, Tb[3] = {0., 0., 0.};
      if (touch) {
        Contact c;
        c.dx = delta[0]; c.dy = delta[1]; c.dz = delta[2];
        c.r = sqrt(rsq); c.rinv = 1.0 / c.r;
        c.radi = xa.w; c.radj = xb.w; c.radsum = xa.w + xb.w; c.deltan_in = 0.0;
        c.mi = va.w; c.mj = vb.w;
        double meff = va.w * vb.w / (va.w + vb.w);
        if (rec_mask(wa.w) & P.freezebit) meff = vb.w;
        if (rec_mask(wb.w) & P.freezebit) meff = va.w;
        c.meff = meff;
        for (int d = 0; d < 3; d++) { c.vi[d] = vav[d]; c.vj[d] = vbv[d]; c.wi[d] = wav[d]; c.wj[d] = wbv[d]; }
        c.itype = ta; c.jtype = tb;
        ContactOut o;
        contact_chain<NORMAL, ROLLING, false>(P, M, c, h, g, su, o, COH == C_BONDNL && bonded);
        for (int d = 0; d < 3; d++) { Fa[d] = o.F[d] + Fb[d]; Ta[d] = o.Ti[d] + Tbi[d]; Tb[d] = o.Tj[d] + Tbj[d]; }
        S = 1.0;
      } else {  // surfacesClose: tangential / rolling history zeroed (tangential_model_history.h:428-440, rolling_model_epsd.h:225-235)
        for (int d = 0; d < 3; d++) { Fa[d] = Fb[d]; Ta[d] = Tbi[d]; Tb[d] = Tbj[d]; h[d] = 0.0; g[d] = 0.0; }
      }
      if (jfirst) { for (int d = 0; d < 3; d++) { F[d] -= Fa[d]; T[d] += Tb[d]; } }
      else { for (int d = 0; d < 3; d++) { F[d] += Fa[d]; T[d] += Ta[d]; } }
      // write the row back: a row is created by the first touch or by a bond creation
      const bool want = had || touch || H[0] != 0.0;
      if (!want) continue;
      if (!had) {
        if (nh < P.hslots) { slot = nh++; P.nbr[(size_t)k * P.lcap + i] = w | ((unsigned)(slot + 1) << NBR_SLOT_SHIFT); }
        else { ((volatile int *)P.flag)[1] = 1; continue; }
      }
      (void)S0;
      {  // rows are always written back: surfacesClose zeroes and the nonlinear trackers move even in the setup step
        double4 *hp = P.hist + (size_t)(slot * M.hrec) * P.lcap + i;
#pragma unroll
        for (int r = 0; r < (COH == C_BOND ? 4 : 8); r++) {
          double4 v;
          v.x = 4 * r < NB ? H[4 * r] : (4 * r == NB ? S : 0.0);
          v.y = 4 * r + 1 < NB ? H[4 * r + 1] : (4 * r + 1 == NB ? S : 0.0);
          v.z = 4 * r + 2 < NB ? H[4 * r + 2] : (4 * r + 2 == NB ? S : 0.0);
          v.w = 4 * r + 3 < NB ? H[4 * r + 3] : (4 * r + 3 == NB ? S : 0.0);
          st4(hp + (size_t)(M.rec_bond + r) * P.lcap, v);
        }
        if (M.tangential) st4(hp + (size_t)M.rec_shear * P.lcap, make_double4(h[0], h[1], h[2], 0.));
        if (HAS_ROLL_HIST) st4(hp + (size_t)M.rec_roll * P.lcap, make_double4(g[0], g[1], g[2], 0.));
      }
    }
    }
    if (nh != nh0) P.numneigh[i] = nn | (nh << 16);
    (void)itype; (void)imask;
    trig = step_epilogue<true>(P, i, xi, vi, wi, F, T);
  }
  if (__any_sync(0xffffffffu, trig) && (threadIdx.x & 31) == 0) *((volatile int *)P.flag) = 1;
}

// ----------------------------------------------------------------------------------------
// Step kernel of the decks whose pair style uses an INL normal law with its own history (hysteretic/nonlinear1|2).  Same
// contract as k_step; one thread per particle; every pair is evaluated by both owners in the canonical orientation (lower tag =
// first body), so the two copies of the 12 + 3 (+3) history values stay bit-identical without a symmetric formulation of the
// law's loading / unloading branches.  contact_distance_factor is 1: entries that do not touch are skipped (history persists
// until the next rebuild drops the pair, pair_gran_base.h:420-424).
template <int NORMAL, int ROLLING>
__global__ void __launch_bounds__(128, 3) k_step_hyst(const StepP P)
{
  constexpr bool HAS_ROLL_HIST = (ROLLING == R_EPSD || ROLLING == R_EPSD2);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool trig = false;
  if (step_gated(P)) return;
  if (i < P.nlocal) {
    const ModelP &M = P.pm;
    const double4 xi = ldg4(P.xr + i), vi = ldg4(P.vm + i), wi = ldg4(P.wt + i);
    const bool su = (P.mode != MODE_SETUP);
    double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
    const int nnw = P.numneigh[i];
    const int nn = nnw & 0xffff;
    int nh = (nnw >> 16) & 0xffff;
    const int nh0 = nh;
    for (int k = 0; k < nn; k++) {
      const unsigned w = P.nbr[(size_t)k * P.lcap + i];
      const int j = (int)(w & NBR_IDX);
      const double4 xj = ldg4(P.xr + j);
      const double dxm = xi.x - xj.x, dym = xi.y - xj.y, dzm = xi.z - xj.z;
      const double rsq = sq3_rn(dxm, dym, dzm);
      const double radsum = xi.w + xj.w;
      if (!(rsq < __dmul_rn(radsum, radsum))) continue;  // pair_gran_base.h:358
      int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
      const bool had = slot >= 0;
      const double4 vj = ldg4(P.vm + j), wj = ldg4(P.wt + j);
      const bool jfirst = (w & NBR_JFIRST) != 0;
      const double4 &xa = jfirst ? xj : xi, &xb = jfirst ? xi : xj, &va = jfirst ? vj : vi, &vb = jfirst ? vi : vj, &wa = jfirst ? wj : wi, &wb = jfirst ? wi : wj;
      double NH[12], h[3] = {0., 0., 0.}, g[3] = {0., 0., 0.}, th0 = 0.0;
#pragma unroll
      for (int d = 0; d < 12; d++) NH[d] = 0.0;
      if (had) {
        const double4 *hp = P.hist + (size_t)(slot * M.hrec) * P.lcap + i;
#pragma unroll
        for (int r = 0; r < 3; r++) { const double4 v = hp[(size_t)(M.rec_norm + r) * P.lcap]; NH[4 * r] = v.x; NH[4 * r + 1] = v.y; NH[4 * r + 2] = v.z; NH[4 * r + 3] = v.w; }
        if (M.tangential) { const double4 v = hp[(size_t)M.rec_shear * P.lcap]; h[0] = v.x; h[1] = v.y; h[2] = v.z; th0 = v.w; }
        if (HAS_ROLL_HIST) { const double4 v = hp[(size_t)M.rec_roll * P.lcap]; g[0] = v.x; g[1] = v.y; g[2] = v.z; }
      }
      Contact c;
      c.dx = xa.x - xb.x; c.dy = xa.y - xb.y; c.dz = xa.z - xb.z;
      c.r = sqrt(rsq); c.rinv = 1.0 / c.r;
      c.radi = xa.w; c.radj = xb.w; c.radsum = xa.w + xb.w; c.deltan_in = 0.0;
      c.mi = va.w; c.mj = vb.w;
      double meff = va.w * vb.w / (va.w + vb.w);
      if (rec_mask(wa.w) & P.freezebit) meff = vb.w;
      if (rec_mask(wb.w) & P.freezebit) meff = va.w;
      c.meff = meff;
      c.vi[0] = va.x; c.vi[1] = va.y; c.vi[2] = va.z; c.vj[0] = vb.x; c.vj[1] = vb.y; c.vj[2] = vb.z;
      c.wi[0] = wa.x; c.wi[1] = wa.y; c.wi[2] = wa.z; c.wj[0] = wb.x; c.wj[1] = wb.y; c.wj[2] = wb.z;
      c.itype = rec_type(wa.w); c.jtype = rec_type(wb.w);
      c.nh = NH; c.th = &th0;
      ContactOut o;
      contact_chain<NORMAL, ROLLING, false>(P, M, c, h, g, su, o);
      if (jfirst) { for (int d = 0; d < 3; d++) { F[d] -= o.F[d]; T[d] += o.Tj[d]; } }
      else { for (int d = 0; d < 3; d++) { F[d] += o.F[d]; T[d] += o.Ti[d]; } }
      if (!had) {
        if (nh < P.hslots) { slot = nh++; P.nbr[(size_t)k * P.lcap + i] = w | ((unsigned)(slot + 1) << NBR_SLOT_SHIFT); }
        else { ((volatile int *)P.flag)[1] = 1; continue; }
      }
      double4 *hp = P.hist + (size_t)(slot * M.hrec) * P.lcap + i;  // the normal law rewrites its values in every evaluation
#pragma unroll
      for (int r = 0; r < 3; r++) st4(hp + (size_t)(M.rec_norm + r) * P.lcap, make_double4(NH[4 * r], NH[4 * r + 1], NH[4 * r + 2], NH[4 * r + 3]));
      if (su || !had || M.tangential == 2) {  // (the hysteretic tangential law restarts its spring in any evaluation)
        if (M.tangential) st4(hp + (size_t)M.rec_shear * P.lcap, make_double4(h[0], h[1], h[2], th0));
        if (HAS_ROLL_HIST) st4(hp + (size_t)M.rec_roll * P.lcap, make_double4(g[0], g[1], g[2], 0.));
      }
    }
    if (nh != nh0) P.numneigh[i] = nn | (nh << 16);
    trig = step_epilogue<true>(P, i, xi, vi, wi, F, T);
  }
  if (__any_sync(0xffffffffu, trig) && (threadIdx.x & 31) == 0) *((volatile int *)P.flag) = 1;
}

// first half step of a run from the stored force arrays: fix_nve_sphere.cpp:134-183
__global__ void __launch_bounds__(256) k_initial_integrate(const StepP P)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool trig = false;
  if (i < P.nlocal) {
    const double4 xi = P.xr[i], vi = P.vm[i], wi = P.wt[i];
    double4 xo = xi, vo = vi, wo = wi;
    if (rec_mask(wi.w) & P.integbit) {
      const double dtfm = P.dtf / vi.w;
      const double dtir = P.dtfrot / (xi.w * xi.w * vi.w);
      vo.x += dtfm * P.f[i]; vo.y += dtfm * P.f[P.cap + i]; vo.z += dtfm * P.f[2 * (size_t)P.cap + i];
      xo.x += P.dtv * vo.x; xo.y += P.dtv * vo.y; xo.z += P.dtv * vo.z;
      wo.x += dtir * P.tq[i]; wo.y += dtir * P.tq[P.cap + i]; wo.z += dtir * P.tq[2 * (size_t)P.cap + i];
    }
    P.xr_o[i] = xo; P.vm_o[i] = vo; P.wt_o[i] = wo;
    const double4 xh = P.xh[i];
    trig = sq3_rn(xo.x - xh.x, xo.y - xh.y, xo.z - xh.z) > P.trigsq;
  }
  if (__any_sync(0xffffffffu, trig) && (threadIdx.x & 31) == 0) *((volatile int *)P.flag) = 1;
}

// ghost refresh of one swap (== CommBrick::forward_comm of x,v,omega for one iswap, comm_brick.cpp:563-645,
// atom_vec_sphere.cpp:283-293): gathers the swap's send list, applies the periodic shift on the sender side
// and writes either straight into the own ghost region (periodic image on the same rank) or into the
// NCCL send buffers.  Lists may contain ghosts received in an earlier dimension (edges/corners).
struct SwapP {
  int n;
  const int *list;
  int dim;
  double shift;
  const double4 *xr, *vm, *wt;
  double4 *ox, *ov, *ow;
};
__global__ void __launch_bounds__(256) k_pack_swap(const SwapP S)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= S.n) return;
  const int s = S.list[q];
  double4 x = S.xr[s];
  if (S.dim == 0) x.x += S.shift; else if (S.dim == 1) x.y += S.shift; else x.z += S.shift;
  S.ox[q] = x; S.ov[q] = S.vm[s]; S.ow[q] = S.wt[s];
}
// Per-step ghost refresh over NVLink peer memory: the sender's pack kernel stores the three records of every send-list
// entry straight into the RECEIVER's ghost region (pointers obtained through CUDA IPC at the last rebuild, see
// halo_p2p_setup) and the last block to finish publishes the exchange's serial number in the receiver's signal word.
// The receiver's stream waits on that word (k_halo_wait) before anything that reads those ghosts.
__global__ void __launch_bounds__(256) k_pack_push(const SwapP S, unsigned *done_counter, volatile int *peer_signal, int serial)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < S.n) {
    const int s = S.list[q];
    double4 x = S.xr[s];
    if (S.dim == 0) x.x += S.shift; else if (S.dim == 1) x.y += S.shift; else x.z += S.shift;
    st4(S.ox + q, x); st4(S.ov + q, S.vm[s]); st4(S.ow + q, S.wt[s]);
  }
  __threadfence_system();  // my peer stores are performed before my block is counted
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) { *done_counter = 0; __threadfence_system(); *peer_signal = serial; }
}
__global__ void k_halo_signal(volatile int *peer_signal, int serial) { *peer_signal = serial; }  // empty send list

// ---- one launch per decomposed dimension and step: both ghost pushes of the dimension and -- in the same launch -- the
// hand-over of this rank's step flags to EVERY rank.  The reference reduces the rebuild decision with MPI_Allreduce
// (neighbor.cpp:1463); here each rank stores its four flag words into its column of every peer's flag box over NVLink peer
// memory and publishes a serial number; k_wait (below) ORs the columns.  No NCCL call between two rebuilds.
struct PushP {
  SwapP S[2]; int nb[2]; unsigned *done[2]; volatile int *sig[2]; int serial[2]; int nsw;
  const int *myflags; int *peer_box[DEM_MAXRANKS]; int me, nranks, slot, fserial, with_flags;
};
__global__ void __launch_bounds__(256) k_push(const PushP Q)
{
  int b = blockIdx.x;
  if (b == (int)gridDim.x - 1) {  // last block: flags to all ranks, and the signal of an empty send list
    const int r = threadIdx.x;
    if (Q.with_flags && r < Q.nranks) {
      int *box = Q.peer_box[r];
#pragma unroll
      for (int k = 0; k < 4; k++) box[(Q.slot * DEM_MAXRANKS + Q.me) * 4 + k] = Q.myflags[k];
      __threadfence_system();
      ((volatile int *)box)[FBOX_SERIAL + Q.me] = Q.fserial;
    }
    if (r >= 32 && r < 32 + Q.nsw && Q.sig[r - 32] && Q.S[r - 32].n == 0) *Q.sig[r - 32] = Q.serial[r - 32];
    return;
  }
  int q = 0;
  if (b >= Q.nb[0]) { q = 1; b -= Q.nb[0]; }
  const SwapP &S = Q.S[q];
  const int e = b * blockDim.x + threadIdx.x;
  if (e < S.n) {
    const int s = S.list[e];
    double4 x = S.xr[s];
    if (S.dim == 0) x.x += S.shift; else if (S.dim == 1) x.y += S.shift; else x.z += S.shift;
    st4(S.ox + e, x); st4(S.ov + e, S.vm[s]); st4(S.ow + e, S.wt[s]);
  }
  __threadfence_system();  // my peer stores are performed before my block is counted
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(Q.done[q], 1u) == (unsigned)Q.nb[q] - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) { *Q.done[q] = 0; __threadfence_system(); *Q.sig[q] = Q.serial[q]; }
}
// receiver side: the ghosts of both swaps have arrived, every rank's flags of this step have arrived; their OR goes to the
// gate words of the next step (device) and, with the serial, to the host's page-locked flag block
struct WaitP {
  const volatile int *sig[2]; int serial[2];
  const int *box; int nranks, slot, fserial, with_flags;
  int *gate_out; int *host_out; volatile int *host_serial; int *timeout_flag;
  int *zero_next;  // the flag slot the NEXT step will write (8 ints), zeroed here instead of by a memset node per step
};
__global__ void k_wait(const WaitP Q)
{  // one warp; every wait is bounded (~4 s): a lost peer must not hang the device
  const int t = threadIdx.x;
  const long long t0 = clock64(), limit = 8000000000LL;
  bool ok = true;
  if (t < 2 && Q.sig[t]) while (*Q.sig[t] < Q.serial[t]) { __nanosleep(64); if (clock64() - t0 > limit) { ok = false; break; } }
  if (Q.with_flags && t < Q.nranks) {
    const volatile int *ser = (const volatile int *)Q.box + FBOX_SERIAL + t;
    while (*ser < Q.fserial) { __nanosleep(64); if (clock64() - t0 > limit) { ok = false; break; } }
  }
  if (!ok) *Q.timeout_flag = 1;
  if (Q.zero_next && t < 8) Q.zero_next[t] = 0;
  __syncwarp();
  __threadfence_system();
  if (Q.with_flags) {
    if (t < 4) {
      int v = 0;
      for (int r = 0; r < Q.nranks; r++) v = max(v, ((const volatile int *)Q.box)[(Q.slot * DEM_MAXRANKS + r) * 4 + t]);
      Q.gate_out[t] = v; Q.host_out[t] = v;
    }
    __syncwarp();
    __threadfence_system();
    if (t == 0) *Q.host_serial = Q.fserial;
  }
}
// Fused ghost push (dem_engine.cu fused_halo_setup): the step kernel has stored the ghost copies from its epilogue and has
// completed (stream order: its peer stores are performed), so ONE warp publishes the exchange's serial number in both
// receivers' signal words, hands this rank's step flags to every rank's flag box, and then waits like k_wait.
struct ShakeP { int *sig_out[2]; int serial_out[2]; const int *myflags; int *peer_box[DEM_MAXRANKS]; int me; WaitP W; };
__global__ void k_handshake(const ShakeP Q)
{
  const int t = threadIdx.x;
  __threadfence_system();
  if (t < 2 && Q.sig_out[t]) *(volatile int *)Q.sig_out[t] = Q.serial_out[t];
  if (t < Q.W.nranks) {
    int *box = Q.peer_box[t];
#pragma unroll
    for (int k = 0; k < 4; k++) box[(Q.W.slot * DEM_MAXRANKS + Q.me) * 4 + k] = Q.myflags[k];
    __threadfence_system();
    ((volatile int *)box)[FBOX_SERIAL + Q.me] = Q.W.fserial;
  }
  // receiver side (== k_wait)
  const WaitP &W = Q.W;
  const long long t0 = clock64(), limit = 8000000000LL;
  bool ok = true;
  if (t < 2 && W.sig[t]) while (*W.sig[t] < W.serial[t]) { __nanosleep(64); if (clock64() - t0 > limit) { ok = false; break; } }
  if (t < W.nranks) {
    const volatile int *ser = (const volatile int *)W.box + FBOX_SERIAL + t;
    while (*ser < W.fserial) { __nanosleep(64); if (clock64() - t0 > limit) { ok = false; break; } }
  }
  if (!ok) *W.timeout_flag = 1;
  if (W.zero_next && t < 8) W.zero_next[t] = 0;
  __syncwarp();
  __threadfence_system();
  if (t < 4) {
    int v = 0;
    for (int r = 0; r < W.nranks; r++) v = max(v, ((const volatile int *)W.box)[(W.slot * DEM_MAXRANKS + r) * 4 + t]);
    W.gate_out[t] = v; W.host_out[t] = v;
  }
  __syncwarp();
  __threadfence_system();
  if (t == 0) *W.host_serial = W.fserial;
}
// single rank: the step's flags to the host's page-locked block (flags, then the serial) and the next step's flag slot
// zeroed -- one tiny launch instead of a memset node, a D2H copy and an event per step
__global__ void k_flags_host(const int *flags, int *host_out, volatile int *host_serial, int serial, int *zero_next)
{
  const int t = threadIdx.x;
  if (t < 4) host_out[t] = ((const volatile int *)flags)[t];
  if (t < 8) zero_next[t] = 0;
  __syncwarp();
  __threadfence_system();
  if (t == 0) *host_serial = serial;
}
__global__ void k_halo_wait(const volatile int *sig_a, int serial_a, const volatile int *sig_b, int serial_b, int *timeout_flag)
{  // bounded (~4 s): a lost peer must not hang the device
  const long long t0 = clock64();
  const long long limit = 8000000000LL;
  if (sig_a) while (*sig_a < serial_a) { __nanosleep(64); if (clock64() - t0 > limit) { *timeout_flag = 1; return; } }
  if (sig_b) while (*sig_b < serial_b) { __nanosleep(64); if (clock64() - t0 > limit) { *timeout_flag = 1; return; } }
  __threadfence_system();
}
// ---- image table of the fused ghost push, built on the device at the end of a rebuild (dem_engine.cu fused_halo_setup).
// gsrc[g] = origin of ghost slot nlocal + g: (kind, a, code) with kind 0: my owned particle a; kind 1 + side: entry a of the
// remote swap of that side as I received it; code = accumulated periodic shifts (2 bits per dimension: +prd, -prd).
// Entries of a particle form a linked list headed by img_first[particle] (order irrelevant: independent stores).
__device__ __forceinline__ void img_link(int *first, int4 *tab, int root, int idx, int slot_t, int code)
{
  tab[idx] = make_int4(slot_t, code, atomicExch(first + root, idx), 0);
}
__global__ void __launch_bounds__(256) k_img_chain(int n, const int *list, int nl, int gfirst, int self, int side, int shiftcode, int4 *gsrc,
                                                   int *rep0, int *rep1, int *cnt)
{  // the new ghosts of one swap (swaps in order); cnt[0..1]: report counters, cnt[2]: unsupported constellation
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int g = gfirst + q - nl;
  if (!self) { gsrc[g] = make_int4(1 + side, q, 0, 0); return; }
  const int sidx = list[q];
  int4 r = sidx < nl ? make_int4(0, sidx, 0, 0) : gsrc[sidx - nl];
  if (((r.z | (r.z >> 1)) & (shiftcode | (shiftcode >> 1)) & 0x15) != 0) cnt[2] = 1;  // two shifts in one dimension
  r.z |= shiftcode;
  gsrc[g] = r;
  if (r.x) {  // a periodic image of a ghost I received: its owner must learn about this copy
    int *rep = r.x == 1 ? rep0 : rep1;
    const int j = atomicAdd(cnt + (r.x - 1), 1);
    rep[3 * j] = r.y; rep[3 * j + 1] = nl + g; rep[3 * j + 2] = r.z;
  }
}
__global__ void __launch_bounds__(256) k_img_local(int ng, int nl, const int4 *gsrc, int *first, int4 *tab)
{
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  const int4 r = gsrc[g];
  if (r.x == 0) img_link(first, tab, r.y, g, nl + g, r.z);
}
__global__ void __launch_bounds__(256) k_img_remote(int n, int direct, const int *rep, const int *list, int nsend, int nl, const int4 *gsrc, int target,
                                                    int pgfirst, int shiftcode, int base, int *first, int4 *tab, int *cnt)
{  // direct: entry k of my send list sits in the receiver's slot pgfirst + k; else: the receiver's reports (entry, slot, code)
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int k = direct ? q : rep[3 * q];
  if (k < 0 || k >= nsend) { cnt[2] = 1; return; }
  const int sidx = list[k];
  int root = sidx, code = 0;
  if (sidx >= nl) { const int4 r = gsrc[sidx - nl]; if (r.x) { cnt[2] = 1; return; } root = r.y; code = r.z; }  // (a ghost forwarded over two rank boundaries: not a slab decomposition)
  const int extra = direct ? 0 : rep[3 * q + 2];
  if ((((code | shiftcode) | ((code | shiftcode) >> 1)) & (extra | (extra >> 1)) & 0x15) != 0 || ((code | (code >> 1)) & (shiftcode | (shiftcode >> 1)) & 0x15) != 0) { cnt[2] = 1; return; }
  img_link(first, tab, root, base + q, (direct ? pgfirst + k : rep[3 * q + 1]) | (target << 28), code | shiftcode | extra);
}
__global__ void __launch_bounds__(256) k_pack_int(int n, const int *list, const int *src, int *out)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = src[list[q]];
}

// ------------------------------------------------------------------ rebuild kernels
struct GridP {
  double org[3], inv[3];
  int nc[3];
  int morton;
};
__device__ __forceinline__ unsigned spread10(unsigned v)
{
  v &= 0x3ff; v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F;
  v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249; return v;
}
__device__ __forceinline__ void cell_of(const GridP &G, const double4 &x, int &cx, int &cy, int &cz)
{
  cx = (int)floor((x.x - G.org[0]) * G.inv[0]); cy = (int)floor((x.y - G.org[1]) * G.inv[1]); cz = (int)floor((x.z - G.org[2]) * G.inv[2]);
  cx = min(max(cx, 0), G.nc[0] - 1); cy = min(max(cy, 0), G.nc[1] - 1); cz = min(max(cz, 0), G.nc[2] - 1);
}
__device__ __forceinline__ int lin_cell(const GridP &G, int cx, int cy, int cz) { return (cz * G.nc[1] + cy) * G.nc[0] + cx; }
__device__ __forceinline__ unsigned key_of(const GridP &G, int cx, int cy, int cz)
{
  return G.morton ? (spread10(cx) | (spread10(cy) << 1) | (spread10(cz) << 2)) : (unsigned)lin_cell(G, cx, cy, cz);
}

// Domain::pbc (domain.cpp:550-640) + sort key of the owned particles
struct BoxP { double lo[3], hi[3], prd[3]; int periodic[3]; };
__global__ void __launch_bounds__(256) k_wrap_key(int n, double4 *xr, const GridP G, const BoxP B, unsigned *keys, int *vals, const int *gone)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gone && gone[i]) { keys[i] = 0xFFFFFFFFu; vals[i] = i; return; }  // migrated away: sorts behind everybody
  double4 x = xr[i];
  {
    double c[3] = {x.x, x.y, x.z};
    bool ch = false;
    for (int d = 0; d < 3; d++) if (B.periodic[d]) {
      if (c[d] < B.lo[d]) { c[d] += B.prd[d]; ch = true; }
      if (c[d] >= B.hi[d]) { c[d] -= B.prd[d]; c[d] = fmax(c[d], B.lo[d]); ch = true; }
    }
    if (ch) { x.x = c[0]; x.y = c[1]; x.z = c[2]; xr[i] = x; }
  }
  int cx, cy, cz; cell_of(G, x, cx, cy, cz);
  keys[i] = key_of(G, cx, cy, cz); vals[i] = i;
}

__global__ void __launch_bounds__(256) k_gather4(int n, const int *perm, const double4 *a, double4 *ao, const double4 *b, double4 *bo,
                                                 const double4 *c, double4 *co)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = perm[i];
  ao[i] = a[p]; bo[i] = b[p]; co[i] = c[p];
}
template <typename T>
__global__ void __launch_bounds__(256) k_gather_rows(int n, int nrows, size_t stride_in, size_t stride_out, const int *perm, const T *a, T *ao)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = perm[i];
  for (int r = 0; r < nrows; r++) ao[r * stride_out + i] = a[r * stride_in + p];
}

// cell ranges of a cell-sorted index range [base, base+n)
__global__ void __launch_bounds__(256) k_cell_ranges(int n, int base, const double4 *xr, const GridP G, int *cstart, int *cend)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int cx, cy, cz;
  cell_of(G, xr[base + p], cx, cy, cz);
  const int c = lin_cell(G, cx, cy, cz);
  int cp = -1, cn = -1;
  if (p > 0) { cell_of(G, xr[base + p - 1], cx, cy, cz); cp = lin_cell(G, cx, cy, cz); }
  if (p < n - 1) { cell_of(G, xr[base + p + 1], cx, cy, cz); cn = lin_cell(G, cx, cy, cz); }
  if (c != cp) cstart[c] = base + p;
  if (c != cn) cend[c] = base + p + 1;
}

// border candidates of one dimension: comm_brick.cpp:884-1117 (slab test; ghosts of earlier dims included)
__global__ void __launch_bounds__(256) k_border_flag(int n, const double4 *xr, int dim, double lo_cut, double hi_cut, int *flo, int *fhi)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double4 x = xr[p];
  const double c = dim == 0 ? x.x : dim == 1 ? x.y : x.z;
  flo[p] = c < lo_cut;    // image at +prd
  fhi[p] = c >= hi_cut;   // image at -prd
}
// compact the flagged indices into a send list (deterministic order: ascending index)
__global__ void __launch_bounds__(256) k_compact(int n, const int *flag, const int *scan, int *list)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n && flag[p]) list[scan[p]] = p;
}
__global__ void __launch_bounds__(256) k_ghost_keys(int n, int nlocal, const double4 *xr, const GridP G, unsigned *keys, int *vals)
{
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  int cx, cy, cz; cell_of(G, xr[nlocal + g], cx, cy, cz);
  keys[g] = key_of(G, cx, cy, cz); vals[g] = nlocal + g;
}
// cell ranges of the ghosts through their cell-sorted index list gorder[0..n)
__global__ void __launch_bounds__(256) k_cell_ranges_idx(int n, const int *gorder, const double4 *xr, const GridP G, int *cstart, int *cend)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int cx, cy, cz;
  cell_of(G, xr[gorder[p]], cx, cy, cz);
  const int c = lin_cell(G, cx, cy, cz);
  int cp = -1, cn = -1;
  if (p > 0) { cell_of(G, xr[gorder[p - 1]], cx, cy, cz); cp = lin_cell(G, cx, cy, cz); }
  if (p < n - 1) { cell_of(G, xr[gorder[p + 1]], cx, cy, cz); cn = lin_cell(G, cx, cy, cz); }
  if (c != cp) cstart[c] = p;
  if (c != cn) cend[c] = p + 1;
}

// ---- particle migration between bricks (CommBrick::exchange, comm_brick.cpp:732-860, with the contact history
// of fix_contact_history.cpp:508-555 travelling with the particle).  One fixed-stride record per migrant:
// [0..11] the three 32-byte records, [12] tag, [13] density, [14] wall-history valid bits, [15] nh,
// [16..16+nwrows) wall history, then the mesh contact rows (mslots x (partner triangle, mhrec x 4 history doubles),
// fix_contact_history_mesh.cpp pack_exchange), then nh x (partner tag, hrec x 4 history doubles).
struct MigP {
  int n, stride, nwrows, hrec, hmax, cap, lcap, dim;
  double wrap_lo, wrap_hi, prd;  // periodic wrap applied by the sender
  int periodic;
  const int *list;
  double4 *xr, *vm, *wt, *xh;
  int *tag; double *density; double *whist;
  unsigned *nbr; int *numneigh; int *ptag; double4 *hist; int hslots, maxk;
  int mslots, mhrec; int *mint; double4 *mhist;  // mesh contact rows (row stride = cap)
  double *buf;
};
__global__ void __launch_bounds__(128) k_mig_pack(const MigP M)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= M.n) return;
  const int i = M.list[q];
  double *b = M.buf + (size_t)q * M.stride;
  double4 x = M.xr[i];
  if (M.periodic) {
    double c = M.dim == 0 ? x.x : M.dim == 1 ? x.y : x.z;
    if (c < M.wrap_lo) c += M.prd;
    if (c >= M.wrap_hi) { c -= M.prd; c = fmax(c, M.wrap_lo); }
    if (M.dim == 0) x.x = c; else if (M.dim == 1) x.y = c; else x.z = c;
  }
  const double4 v = M.vm[i], w = M.wt[i];
  b[0] = x.x; b[1] = x.y; b[2] = x.z; b[3] = x.w; b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w;
  b[8] = w.x; b[9] = w.y; b[10] = w.z; b[11] = w.w;
  b[12] = (double)M.tag[i]; b[13] = M.density[i];
  b[14] = (double)(((unsigned)(__double_as_longlong(M.xh[i].w) & 0xffffffffLL)) >> 16);
  for (int r = 0; r < M.nwrows; r++) b[16 + r] = M.whist[(size_t)r * M.cap + i];
  const int mblock = M.mslots * (1 + 4 * M.mhrec);
  for (int s = 0; s < M.mslots; s++) {
    double *e = b + 16 + M.nwrows + (size_t)s * (1 + 4 * M.mhrec);
    e[0] = (double)M.mint[(size_t)(1 + s) * M.cap + i];
    for (int r = 0; r < M.mhrec; r++) {
      const double4 h = M.mhist[(size_t)(s * M.mhrec + r) * M.cap + i];
      e[1 + 4 * r] = h.x; e[2 + 4 * r] = h.y; e[3 + 4 * r] = h.z; e[4 + 4 * r] = h.w;
    }
  }
  int nh = 0;
  if (M.nbr) {
    const int nn = M.numneigh[i] & 0xffff;
    double *hb = b + 16 + M.nwrows + mblock;
    for (int k = 0; k < nn; k++) {
      const unsigned wd = M.nbr[(size_t)k * M.lcap + i];
      if (!(wd & NBR_HIST) || nh >= M.hmax) continue;
      const int slot = (int)((wd & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
      double *e = hb + (size_t)nh * (1 + 4 * M.hrec);
      e[0] = (double)M.ptag[(size_t)k * M.lcap + i];
      for (int r = 0; r < M.hrec; r++) {
        const double4 h = M.hist[(size_t)(slot * M.hrec + r) * M.lcap + i];
        e[1 + 4 * r] = h.x; e[2 + 4 * r] = h.y; e[3 + 4 * r] = h.z; e[4 + 4 * r] = h.w;
      }
      nh++;
    }
  }
  b[15] = (double)nh;
}
// arrivals are appended behind the current particles at index base+q; their history goes into row base+q of
// the OLD list so that the remap of k_build_list finds it through the sort permutation like any other row
__global__ void __launch_bounds__(128) k_mig_unpack(const MigP M, int base)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= M.n) return;
  const int i = base + q;
  const double *b = M.buf + (size_t)q * M.stride;
  M.xr[i] = make_double4(b[0], b[1], b[2], b[3]);
  M.vm[i] = make_double4(b[4], b[5], b[6], b[7]);
  M.wt[i] = make_double4(b[8], b[9], b[10], b[11]);
  M.tag[i] = (int)b[12]; M.density[i] = b[13];
  M.xh[i] = make_double4(0., 0., 0., __longlong_as_double((long long)(((unsigned)b[14]) << 16)));
  for (int r = 0; r < M.nwrows; r++) M.whist[(size_t)r * M.cap + i] = b[16 + r];
  const int mblock = M.mslots * (1 + 4 * M.mhrec);
  if (M.mslots) M.mint[i] = 0;  // candidate count: rebuilt by k_mesh_cand
  for (int s = 0; s < M.mslots; s++) {
    const double *e = b + 16 + M.nwrows + (size_t)s * (1 + 4 * M.mhrec);
    M.mint[(size_t)(1 + s) * M.cap + i] = (int)e[0];
    for (int r = 0; r < M.mhrec; r++) M.mhist[(size_t)(s * M.mhrec + r) * M.cap + i] = make_double4(e[1 + 4 * r], e[2 + 4 * r], e[3 + 4 * r], e[4 + 4 * r]);
  }
  if (M.nbr) {
    const int nh = (int)b[15];
    const double *hb = b + 16 + M.nwrows + mblock;
    for (int k = 0; k < nh; k++) {
      const double *e = hb + (size_t)k * (1 + 4 * M.hrec);
      M.ptag[(size_t)k * M.lcap + i] = (int)e[0];
      M.nbr[(size_t)k * M.lcap + i] = (unsigned)(k + 1) << NBR_SLOT_SHIFT;
      for (int r = 0; r < M.hrec; r++)
        M.hist[(size_t)(k * M.hrec + r) * M.lcap + i] = make_double4(e[1 + 4 * r], e[2 + 4 * r], e[3 + 4 * r], e[4 + 4 * r]);
    }
    M.numneigh[i] = nh | (nh << 16);
  }
}
__global__ void __launch_bounds__(256) k_mig_flag(int n, const double4 *xr, int dim, double sublo, double subhi, int has_lo, int has_hi,
                                                  int *flo, int *fhi, int *gone)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (gone[p]) { flo[p] = fhi[p] = 0; return; }
  const double4 x = xr[p];
  const double c = dim == 0 ? x.x : dim == 1 ? x.y : x.z;
  const int l = has_lo && c < sublo, h = has_hi && c >= subhi;
  flo[p] = l; fhi[p] = h;
  if (l || h) gone[p] = 1;
}
__global__ void __launch_bounds__(256) k_max_nh(int n, const int *list, const int *numneigh, int *out)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) atomicMax(out, (numneigh[list[q]] >> 16) & 0xffff);
}

// Verlet-skin FULL list + history remap: neigh_gran.cpp:560-625 (predicates), fix_contact_history.cpp:351
struct BuildP {
  int nlocal, cap, maxk, dnum, hslots;
  const double4 *xr;
  const int *tag;
  GridP G;
  const int *ocs, *oce, *gcs, *gce;  // owned / ghost cell ranges
  const int *gorder;                 // cell-sorted ghost indices (ghost storage keeps its swap order)
  double cdf, skin;
  unsigned *nbr; int *numneigh; int *ptag; double4 *hist;
  // previous list (rows addressed through perm: new i <- old perm[i])
  int have_old, cap_old, dnum_old;
  int coh_nbond, coh_rec;  // bond models: a row is kept iff its bondFlag or its sticky flag is set (see k_step_bond)
  const int *perm; const unsigned *nbr_old; const int *numneigh_old; const int *ptag_old; const double4 *hist_old;
  int *overflow;
};
__global__ void __launch_bounds__(128) k_build_list(const BuildP B)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.nlocal) return;
  const double4 xi = B.xr[i];
  const int tagi = B.tag[i];
  int cx, cy, cz; cell_of(B.G, xi, cx, cy, cz);
  const int oi = B.have_old ? B.perm[i] : 0;
  const int nold = B.have_old ? (B.numneigh_old[oi] & 0xffff) : 0;
  int n = 0, nh = 0, nband = 0;  // list entries, kept history rows, entries inside the contact band right now
  for (int dz = -1; dz <= 1; dz++) {
    const int z = cz + dz; if (z < 0 || z >= B.G.nc[2]) continue;
    for (int dy = -1; dy <= 1; dy++) {
      const int y = cy + dy; if (y < 0 || y >= B.G.nc[1]) continue;
      for (int dx = -1; dx <= 1; dx++) {
        const int x = cx + dx; if (x < 0 || x >= B.G.nc[0]) continue;
        const int c = lin_cell(B.G, x, y, z);
        for (int pass = 0; pass < 2; pass++) {
          const int s = pass ? B.gcs[c] : B.ocs[c], e = pass ? B.gce[c] : B.oce[c];
          for (int q = s; q < e; q++) {
            const int j = pass ? B.gorder[q] : q;
            if (j == i) continue;
            const double4 xj = B.xr[j];
            const double rsq = sq3_rn(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
            const double radsum = __dmul_rn(xi.w + xj.w, B.cdf);
            const double rc = radsum + B.skin;
            if (rsq <= __dmul_rn(rc, rc)) {
              if (!B.coh_nbond && rsq < __dmul_rn(radsum, radsum)) nband++;  // (bond decks: the band is wide, rows follow the bonds)
              if (n < B.maxk) {
                const int tagj = B.tag[j];
                unsigned w = (unsigned)j | (tagj < tagi ? NBR_JFIRST : 0u);
                if (nold && rsq < __dmul_rn(radsum, radsum)) w |= NBR_HIST;  // inside the contact band: candidate for a kept history row (resolved below)
                B.nbr[(size_t)n * B.cap + i] = w;
                B.ptag[(size_t)n * B.cap + i] = tagj;
              }
              n++;
            }
          }
        }
      }
    }
  }
  // history remap, after the stencil walk and with all lanes on the same row entry: inside the walk a lane's in-band hits
  // come at different candidates than its neighbours', so the search below ran in two thirds of the ~150 candidate
  // iterations of a warp instead of once per row entry (4.3 of the 6.5 ms of a rebuild of the 4.19M bed)
  if (nold) {
    const int nrow = min(n, B.maxk);
    for (int k = 0; k < nrow; k++) {
      unsigned w = B.nbr[(size_t)k * B.cap + i];
      if ((w & NBR_HIST) != NBR_HIST) continue;
      w &= ~NBR_HIST;
      const int tagj = B.ptag[(size_t)k * B.cap + i];
      // (a settled bed: the rows hardly change between two rebuilds, so the old row is searched from this entry's own position)
      for (int mm = 0; mm < nold; mm++) {
        int m = (k < nold ? k : 0) + mm; if (m >= nold) m -= nold;  // k, k+1, .., wrapping: a rotation of 0..nold-1
        const unsigned wo = B.nbr_old[(size_t)m * B.cap_old + oi];
        if ((wo & NBR_HIST) && B.ptag_old[(size_t)m * B.cap_old + oi] == tagj) {
          const int so = (int)((wo & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
          const int srec = B.coh_rec + B.coh_nbond / 4, scomp = B.coh_nbond % 4;
          if (B.coh_nbond) {  // reference: contact_flags == 0 rows are not partners (fix_contact_history.cpp:351)
            const double4 b0 = B.hist_old[(size_t)(so * B.dnum + B.coh_rec) * B.cap_old + oi], sr = B.hist_old[(size_t)(so * B.dnum + srec) * B.cap_old + oi];
            const double S = scomp == 0 ? sr.x : scomp == 1 ? sr.y : scomp == 2 ? sr.z : sr.w;
            if (b0.x == 0.0 && S == 0.0) break;
          }
          if (nh < B.hslots) {
            w |= (unsigned)(nh + 1) << NBR_SLOT_SHIFT;
            for (int d = 0; d < B.dnum; d++) {  // dnum = 32-byte records per contact here
              double4 v = B.hist_old[(size_t)(so * B.dnum + d) * B.cap_old + oi];
              if (B.coh_nbond && d == srec) { if (scomp == 0) v.x = 1.0; else if (scomp == 1) v.y = 1.0; else if (scomp == 2) v.z = 1.0; else v.w = 1.0; }  // kept rows restart with flag 1
              B.hist[(size_t)(nh * B.dnum + d) * B.cap + i] = v;
            }
          }
          nh++;
          break;
        }
      }
      B.nbr[(size_t)k * B.cap + i] = w;
    }
  }
  B.numneigh[i] = min(n, B.maxk) | (min(nh, B.hslots) << 16);
  if (n > B.maxk) atomicMax(B.overflow, n);
  // history slots: the rows kept from the old list, or -- a bed uploaded in a packed state, an overlapping initial
  // condition -- every entry that is inside the contact band now and will ask for a row in the first step, plus 8 spare
  if (max(nh, nband) + 8 > B.hslots) atomicMax(B.overflow + 1, max(nh, nband));
}

// positions at build time (neighbor.cpp:1486-1510) + primitive-wall candidate bits (primitive_wall.h:129-138)
__global__ void __launch_bounds__(256) k_wall_index(int n, const int *flag, const int *scan, int *wlist, double4 *xh)
{  // compact list of wall candidates + the particle's position in it (upper 32 bits of xh.w, +1)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const int c = scan[i];
  wlist[c] = i;
  const long long b = (__double_as_longlong(xh[i].w) & 0xffffffffLL) | ((long long)(c + 1) << 32);
  xh[i].w = __longlong_as_double(b);
}
__global__ void __launch_bounds__(256) k_hold(int n, const double4 *xr, double4 *xh, const unsigned *valid_in, const WallP *walls, int nwalls, double skin, int *cflag,
                                              const int *mesh_ncand)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 x = xr[i];
  unsigned cand = 0;
  const double pos[3] = {x.x, x.y, x.z};
  for (int w = 0; w < nwalls; w++) {
    const WallP &W = walls[w];
    bool in;
    const double dMax = x.w + skin;
    if (W.wtype < 3) {
      const double dist = pos[W.wtype] - W.param[0];
      in = ((dist > 0.0) ? dist : -dist) <= dMax;
    } else {
      const int dd = W.wtype - 3;
      const double dy = pos[(dd + 1) % 3] - W.param[1], dz = pos[(dd + 2) % 3] - W.param[2];
      const double dist = sqrt(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz))) - W.param[0];
      in = (dMax < dist || -dMax < dist);
    }
    if (in) cand |= 1u << w;
  }
  const unsigned valid = valid_in ? valid_in[i] : 0u;
  double4 o; o.x = x.x; o.y = x.y; o.z = x.z; o.w = __longlong_as_double((long long)(cand | (valid << 16)));
  xh[i] = o;
  if (cflag) cflag[i] = cand != 0 || (mesh_ncand && mesh_ncand[i] > 0);  // compact wall list: primitive or mesh candidates
}
__global__ void __launch_bounds__(256) k_extract_valid(int n, const int *perm, const double4 *xh, unsigned *valid)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  valid[i] = ((unsigned)(__double_as_longlong(xh[perm ? perm[i] : i].w) & 0xffffffffLL)) >> 16;
}

// mass of a sphere created by fix insert/*: density * volume with the volume rounded the way FixTemplateSphere::randomize_ptilist
// forms it (fix_template_sphere.cpp:349-350), which is not the rounding of read_data / set (atom_vec_sphere.cpp:1078)
__device__ __forceinline__ double insert_mass(double r, double rho)
{
  const double vol = r * r * r * 4. * 3.14159265358979323846 / 3.;
  return rho * vol;
}
// host-array upload: builds the three 32-byte records on the device from the caller's plain arrays
// (atom_vec_sphere.cpp:1055-1083: radius = diameter/2 is done by the caller, rmass = 4/3 pi r^3 rho here)
__global__ void __launch_bounds__(256) k_pack_upload(int n, const double *x, const double *v, const double *omega, const double *radius,
                                                     const double *density, const int *type, const int *mask, const int *tag, int ntypes,
                                                     double4 *xr, double4 *vm, double4 *wt, int *err, unsigned long long *rmax_bits, int ins_mass)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long b = 0ull, bm = ~0ull;  // running maximum / minimum of the radius bit patterns (rmax_bits[0] / rmax_bits[2])
  if (i < n) {
    const double r = radius[i], rho = density[i];
    const int t = type[i];
    if (t < 1 || t > ntypes) atomicOr(err, 1);
    if (!(r > 0.0) || !(rho > 0.0)) atomicOr(err, 2);
    if (tag[i] <= 0) atomicOr(err, 4);
    const double m = ins_mass ? insert_mass(r, rho) : 4.0 * 3.14159265358979323846 / 3.0 * r * r * r * rho;
    xr[i] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], r);
    vm[i] = make_double4(v ? v[3 * i] : 0., v ? v[3 * i + 1] : 0., v ? v[3 * i + 2] : 0., m);
    wt[i] = make_double4(omega ? omega[3 * i] : 0., omega ? omega[3 * i + 1] : 0., omega ? omega[3 * i + 2] : 0.,
                         __longlong_as_double(pack_bits(t, mask ? mask[i] : 1)));
    b = (unsigned long long)__double_as_longlong(r > 0.0 ? r : 0.0); bm = b;
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long ob = __shfl_down_sync(0xffffffffu, b, o); b = ob > b ? ob : b;
    const unsigned long long om = __shfl_down_sync(0xffffffffu, bm, o); bm = om < bm ? om : bm;
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(rmax_bits, b); atomicMin(rmax_bits + 2, bm); }  // positive doubles order like their bit patterns
}
// multi-rank upload: every rank receives the whole particle set and keeps its brick.  Ownership test on the wrapped
// position (Domain::pbc + sub-box, like read_data.cpp / atom.cpp data_atoms); validity checks and the global maximum
// radius run over ALL particles.  flag[i] = 1 when the particle is mine.
struct MineP { double lo[3], hi[3], prd[3], sublo[3], subhi[3]; int periodic[3], first[3], last[3]; };
__host__ __device__ __forceinline__ bool brick_owns(const MineP &B, const double *x3)
{
  bool mine = true;
  for (int d = 0; d < 3 && mine; d++) {
    double c = x3[d];
    if (B.periodic[d]) { if (c < B.lo[d]) c += B.prd[d]; if (c >= B.hi[d]) { c -= B.prd[d]; c = c > B.lo[d] ? c : B.lo[d]; } }
    const bool lo_ok = c >= B.sublo[d] || (B.first[d] && !B.periodic[d]);
    const bool hi_ok = c < B.subhi[d] || (B.last[d] && !B.periodic[d]);
    mine = lo_ok && hi_ok;
  }
  return mine;
}
__global__ void __launch_bounds__(256) k_flag_mine(int n, const double *x, const double *radius, const double *density, const int *type, const int *tag,
                                                   int ntypes, const MineP B, int *flag, int *err, unsigned long long *rmax_bits)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long b = 0ull, bm = ~0ull;
  if (i < n) {
    const double r = radius[i], rho = density[i];
    const int t = type[i];
    if (t < 1 || t > ntypes) atomicOr(err, 1);
    if (!(r > 0.0) || !(rho > 0.0)) atomicOr(err, 2);
    if (tag[i] <= 0) atomicOr(err, 4);
    flag[i] = brick_owns(B, x + 3 * (size_t)i) ? 1 : 0;
    b = (unsigned long long)__double_as_longlong(r > 0.0 ? r : 0.0); bm = b;
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long ob = __shfl_down_sync(0xffffffffu, b, o); b = ob > b ? ob : b;
    const unsigned long long om = __shfl_down_sync(0xffffffffu, bm, o); bm = om < bm ? om : bm;
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(rmax_bits, b); atomicMin(rmax_bits + 2, bm); }
}
// records of the selected particles (ascending original index, like the host loop it replaces)
__global__ void __launch_bounds__(256) k_pack_upload_sel(int nsel, const int *list, const double *x, const double *v, const double *omega, const double *radius,
                                                         const double *density, const int *type, const int *mask, const int *tag,
                                                         double4 *xr, double4 *vm, double4 *wt, int *otag, double *odensity, int ins_mass)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nsel) return;
  const int i = list[q];
  const double r = radius[i], rho = density[i];
  const double m = ins_mass ? insert_mass(r, rho) : 4.0 * 3.14159265358979323846 / 3.0 * r * r * r * rho;  // atom_vec_sphere.cpp:1078
  xr[q] = make_double4(x[3 * i], x[3 * i + 1], x[3 * i + 2], r);
  vm[q] = make_double4(v ? v[3 * i] : 0., v ? v[3 * i + 1] : 0., v ? v[3 * i + 2] : 0., m);
  wt[q] = make_double4(omega ? omega[3 * i] : 0., omega ? omega[3 * i + 1] : 0., omega ? omega[3 * i + 2] : 0.,
                       __longlong_as_double(pack_bits(type[i], mask ? mask[i] : 1)));
  otag[q] = tag[i]; odensity[q] = rho;
}
// read-back in ascending tag order: field 0..2 = xyz of a record array, 3 = .w, 4 = type, 5 = mask
__global__ void __launch_bounds__(256) k_gather_out(int n, const int *order, const double4 *rec, int what, double *outd, int *outi)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double4 r = rec[order[k]];
  if (what == 0) { outd[3 * k] = r.x; outd[3 * k + 1] = r.y; outd[3 * k + 2] = r.z; }
  else if (what == 3) outd[k] = r.w;
  else if (what == 4) outi[k] = rec_type(r.w);
  else outi[k] = rec_mask(r.w);
}
__global__ void __launch_bounds__(256) k_gather_rows_out(int n, const int *order, const double *src, size_t stride, int rows, double *out)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  for (int r = 0; r < rows; r++) out[(size_t)rows * k + r] = src[r * stride + order[k]];
}
__global__ void __launch_bounds__(256) k_iota(int n, int *v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }

__global__ void __launch_bounds__(256) k_count_pairs(int n, const int *numneigh, const unsigned *nbr, int cap, unsigned long long *out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a = 0, b = 0;
  if (i < n) { const int nn = numneigh[i] & 0xffff; a = nn; for (int k = 0; k < nn; k++) b += (nbr[(size_t)k * cap + i] & NBR_HIST) ? 1 : 0; }
  for (int o = 16; o; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { if (a) atomicAdd(out, a); if (b) atomicAdd(out + 1, b); }
}

}  // namespace dem
