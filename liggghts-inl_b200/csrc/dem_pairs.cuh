// dem_pairs.cuh -- the pair step of the plain contact models (no cohesion), second generation ("owner list").
//
// Why: the first-generation k_step evaluated every contact twice (once per owner of a FULL list) and kept two copies of
// its history; ncu showed the kernel issue/latency bound with a third of its 561 M warp instructions in fp64 and 1.9x the
// algorithmic DRAM traffic (profiles/r02_*).  Here every pair has ONE owner -- the particle with the lower storage index,
// or the local particle when the partner is a ghost -- which evaluates the contact once, keeps the single copy of its
// history and leaves the partner's share (-F, torque on the partner) in a per-contact result record.  The partner picks
// the record up through its own row.  Nothing is accumulated with atomics: a particle's sum always runs over its row in
// row order (own contacts first, then the received shares), so runs stay bit reproducible -- what the reference gets from
// its half list with `newton off` (pair_gran_base.h:257-496, f[j] updated by the owner of the pair) without the races.
//
// Row layout (ELLPACK, transposed, stride lcap; numneigh word = NN2_PACK(total, owned, history slots in use)):
//   entries [0, owned)              pairs this particle owns: partner index > own index, or partner is a ghost
//   entries [maxk - (total-owned), maxk)  pairs the partner owns, filled from the back of the row
// A neighbour word is [31] partner tag < own tag, [30:25] history slot + 1 of the pair IN THE OWNER'S ROW (0: the pair has
// no history == reference contact_flag 0), [24:0] partner index.  History records live at hist[(slot*hrec + r)*lcap + owner],
// result records in a ring of chunk-sized buffers (res_at below; one aligned 64-byte block per pair): h=0 (Fx, Fy, Fz, Tpx), h=1 (Tpy, Tpz, serial, owner index) with F the force
// on the owner and Tp the torque on the partner; `serial` = the launch that wrote it (a partner ignores stale records and
// records of another chunk that shares the ring slot).
#pragma once
#include "dem_kernels.cuh"

namespace dem {

#define NN2_TOT(w) ((int)((w) & 0x3ff))
#define NN2_OWN(w) ((int)(((w) >> 10) & 0x3ff))
#define NN2_NH(w) ((int)(((w) >> 20) & 0x3f))
#define NN2_PACK(tot, own, nh) ((int)((tot) | ((own) << 10) | ((nh) << 20)))
#define NN2_MAXK 1023

__device__ __forceinline__ unsigned ldcg_u32(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double4 ldcg4(const double4 *p)
{  // L2 read (results another SM wrote in this launch must not come from a stale L1 line)
  double4 v;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

// L2 eviction hints (createpolicy + .L2::cache_hint): the rings must stay in L2 from the moment a record is written until
// the slot is overwritten a few chunks later, while ~1 KB per particle of use-once data (rows, history, fresh records)
// streams through the same cache.  Ring traffic is marked evict_last, the use-once streams evict_first.
#ifndef DEM_L2HINTS
#define DEM_L2HINTS 1
#endif
struct L2Pol { unsigned long long first, last; };
__device__ __forceinline__ L2Pol l2_policies()
{
  L2Pol p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.first));
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.last));
  return p;
}
__device__ __forceinline__ double4 ld4_pol(const double4 *p, unsigned long long pol)
{
  double4 v;
#if DEM_L2HINTS
  asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol));
#else
  v = *p;
#endif
  return v;
}
__device__ __forceinline__ double4 ldcg4_pol(const double4 *p, unsigned long long pol)
{
  double4 v;
#if DEM_L2HINTS
  asm volatile("ld.global.cg.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p), "l"(pol));
#else
  v = ldcg4(p);
#endif
  return v;
}
__device__ __forceinline__ void st4_pol(double4 *p, const double4 &v, unsigned long long pol)
{
#if DEM_L2HINTS
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w), "l"(pol));
#else
  st4(p, v);
#endif
}
__device__ __forceinline__ double ldcg_pol(const double *p, unsigned long long pol)
{
#if DEM_L2HINTS
  double v; asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); return v;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void st_pol(double *p, double v, unsigned long long pol)
{
#if DEM_L2HINTS
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol));
#else
  *p = v;
#endif
}
__device__ __forceinline__ unsigned ldu32_pol(const unsigned *p, unsigned long long pol)
{
#if DEM_L2HINTS
  unsigned v; asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v;
#else
  return *p;
#endif
}
__device__ __forceinline__ unsigned ldcgu32_pol(const unsigned *p, unsigned long long pol)
{
#if DEM_L2HINTS
  unsigned v; asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v;
#else
  return ldcg_u32(p);
#endif
}

// A step runs as a wavefront over chunks (= slabs of the storage order, dem_kernels.cuh GridP): phase A of chunk c (owned
// pairs) must be complete before phase B (shares of the partner-owned pairs + integration) of chunks c and c+1 runs --
// the owner of a pair has the lower index, so it sits in the same or the previous chunk.  Results and partial sums of a
// chunk live in slot (c % ring) of a ring: the addresses are re-used a few chunks later, while the lines are still in
// L2, so this traffic never reaches DRAM.
struct Chunk {
  int c, cs, ce;             // chunk index, first particle, one past the last
  int cs_prev;               // first particle of the previous chunk
  double4 *res, *res_prev;   // ring buffers of result records of this / the previous chunk
  double *part;              // ring buffer of partial sums of this chunk
  L2Pol pol;
};
__device__ __forceinline__ Chunk chunk_of(const StepP &P, int c, int cs, int ce, int cs_prev)
{
  Chunk K;
  K.c = c; K.cs = cs; K.ce = ce; K.cs_prev = cs_prev;
  const int r = c % P.ring, rp = (c + P.ring - 1) % P.ring;
  K.res = P.res + (size_t)r * P.hslots * P.ccap * 2; K.res_prev = P.res + (size_t)rp * P.hslots * P.ccap * 2;
  K.part = P.part + (size_t)r * 6 * P.ccap;
  K.pol = l2_policies();
  return K;
}
__device__ __forceinline__ double4 *res_rec(const StepP &P, double4 *base, int slot, int off) { return base + ((size_t)slot * P.ccap + off) * 2; }

// one owned, touching pair of particle i: evaluated once, in the owner's orientation (pair_chain); history in the canonical
// orientation "lower tag first"; the partner's share goes to the pair's result record
#ifndef DEM_P_INLINE
#define DEM_P_INLINE __forceinline__
#endif
template <int NORMAL, int ROLLING, bool ONE>
__device__ DEM_P_INLINE void pair_contact2(const StepP &P, const Chunk &K, int i, unsigned w, const double4 &xi, const double4 &vi, const double4 &wi,
                                              bool su, int *nh, double *F, double *T)
{
  constexpr bool HAS_ROLL_HIST = (ROLLING == R_EPSD || ROLLING == R_EPSD2);
  const int j = (int)(w & NBR_IDX);
  int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
  const bool had = slot >= 0;
  const double4 xj = ldg4(P.xr + j), vj = ldg4(P.vm + j), wj = ldg4(P.wt + j);
  double4 hs = make_double4(0., 0., 0., 0.), hr = make_double4(0., 0., 0., 0.);
  if (had) {
    const double4 *hp = P.hist + (size_t)(slot * P.pm.hrec) * P.lcap + i;
    if (P.pm.tangential) hs = ld4_pol(hp + (size_t)P.pm.rec_shear * P.lcap, K.pol.first);
    if (HAS_ROLL_HIST) hr = ld4_pol(hp + (size_t)P.pm.rec_roll * P.lcap, K.pol.first);
  }
  const double sgn = (w & NBR_JFIRST) ? -1.0 : 1.0;
  double h[3] = {sgn * hs.x, sgn * hs.y, sgn * hs.z}, g[3] = {sgn * hr.x, sgn * hr.y, sgn * hr.z};
  const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
  const double rsq = sq3_rn(dx, dy, dz);
  double Fc[3] = {0., 0., 0.}, Tc[3] = {0., 0., 0.}, Tp[3] = {0., 0., 0.};
  pair_chain<NORMAL, ROLLING, ONE>(P, P.pm, xi, vi, wi, xj, vj, wj, rec_type(wi.w), rec_type(wj.w), rec_mask(wi.w), rec_mask(wj.w), dx, dy, dz, rsq, h, g, su, Fc, Tc, Tp);
#pragma unroll
  for (int d = 0; d < 3; d++) { F[d] += Fc[d]; T[d] += Tc[d]; }
  if (!had) {  // first touch since the last rebuild: the contact flag becomes != 0 and stays; both rows learn the slot
    const int s = atomicAdd(nh, 1);
    if (s < P.hslots) {
      slot = s;
      const unsigned wn = w | ((unsigned)(slot + 1) << NBR_SLOT_SHIFT);
      const int nown = NN2_OWN(P.numneigh[i]);
      for (int k = 0; k < nown; k++)  // rare path: the row entry of this partner
        if ((P.nbr[(size_t)k * P.lcap + i] & NBR_IDX) == (unsigned)j) { P.nbr[(size_t)k * P.lcap + i] = wn; break; }
      if (j < P.nlocal) {  // ... and the partner's entry for me, in the back region of its row
        const int nnw = P.numneigh[j];
        const int nnon = NN2_TOT(nnw) - NN2_OWN(nnw);
        for (int m = 0; m < nnon; m++) {
          unsigned *q = P.nbr + (size_t)(P.maxk - 1 - m) * P.lcap + j;
          const unsigned wq = *q;
          if ((wq & NBR_IDX) == (unsigned)i) { *q = (wq & ~NBR_HIST) | ((unsigned)(slot + 1) << NBR_SLOT_SHIFT); break; }
        }
      }
    } else { atomicSub(nh, 1); ((volatile int *)P.flag)[1] = 1; }
  }
  if (slot >= 0) {
    if (P.pm.hrec && (su || !had)) {
      double4 *hp = P.hist + (size_t)(slot * P.pm.hrec) * P.lcap + i;
      if (P.pm.tangential) st4_pol(hp + (size_t)P.pm.rec_shear * P.lcap, make_double4(sgn * h[0], sgn * h[1], sgn * h[2], 0.), K.pol.first);
      if (HAS_ROLL_HIST) st4_pol(hp + (size_t)P.pm.rec_roll * P.lcap, make_double4(sgn * g[0], sgn * g[1], sgn * g[2], 0.), K.pol.first);
    }
    if (j < P.nlocal) {
      double4 *rp = res_rec(P, K.res, slot, i - K.cs);
      st4_pol(rp, make_double4(Fc[0], Fc[1], Fc[2], Tp[0]), K.pol.last);
      st4_pol(rp + 1, make_double4(Tp[1], Tp[2], P.serial, (double)i), K.pol.last);
    }
  } else if (j < P.nlocal) ((volatile int *)P.flag)[1] = 1;  // no slot: the partner cannot be served (reported as history overflow)
}

__device__ __forceinline__ void prefetch_contact2(const StepP &P, int i, unsigned w)
{
  const int j = (int)(w & NBR_IDX);
  prefetch_l2(P.vm + j); prefetch_l2(P.wt + j);
  const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
  if (slot >= 0) {
    const double4 *hp = P.hist + (size_t)(slot * P.pm.hrec) * P.lcap + i;
    for (int r = 0; r < P.pm.hrec; r++) prefetch_l2(hp + (size_t)r * P.lcap);
  }
}

#ifndef DEM_P_CMAX
#define DEM_P_CMAX 8   // owned touching entries per particle staged in shared memory (more: evaluated on the spot)
#endif
#ifndef DEM_P_OWNR
#define DEM_P_OWNR 8
#endif
#ifndef DEM_P_SWEEPW
#define DEM_P_SWEEPW 5
#endif
#ifndef DEM_P_MINBLOCKS
#define DEM_P_MINBLOCKS 5
#endif
#ifndef DEM_P_WAVE_PREFETCH
#define DEM_P_WAVE_PREFETCH 200
#endif

// Phase A of a step, for the 128 particles [i0, i0+128): the owned pairs.  Same warp machinery as the first generation
// (sweep in passes with the gathers in flight, own-lane rounds, cooperative deal of the uneven remainder), half the entries.
// Leaves the particle's partial sum (its owned contacts, row order) in P.f / P.tq.
template <int NORMAL, int ROLLING, bool ONE>
__device__ __forceinline__ void pairs_block(const StepP &P, const Chunk &K, int i0, unsigned (*s_w)[128], double4 (*s_rec)[128], double (*s_res)[4 * 32], int *s_off, int *s_nh)
{
  const int tid = threadIdx.x, lane = tid & 31, wb = tid & ~31;
  const int i = i0 + tid;
  const bool active = i < K.ce;
  const bool su = (P.mode != MODE_SETUP);
  double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
  int nc = 0, nh0 = 0, nown = 0, nnw = 0;
#if DEM_P_WAVE_PREFETCH > 0
  {
    const int ip = i + DEM_P_WAVE_PREFETCH * 128;
    if (ip < P.nlocal) {
      prefetch_l2(P.xr + ip); prefetch_l2(P.vm + ip); prefetch_l2(P.wt + ip);
      if (lane < 8) prefetch_l2(P.nbr + (size_t)lane * P.lcap + (ip - lane));
      else if (lane == 8) prefetch_l2(P.numneigh + (ip - lane));
    }
  }
#endif
  {
    double4 xi = make_double4(0., 0., 0., 0.), vi = xi, wi = xi;
    if (active) { xi = ldg4(P.xr + i); vi = ldg4(P.vm + i); wi = ldg4(P.wt + i); }
    s_rec[0][tid] = xi; s_rec[1][tid] = vi; s_rec[2][tid] = wi;
    if (active) { nnw = P.numneigh[i]; nown = NN2_OWN(nnw); nh0 = NN2_NH(nnw); }
    s_nh[tid] = nh0;
    for (int k0 = 0; k0 < nown; k0 += DEM_P_SWEEPW) {
      unsigned wv[DEM_P_SWEEPW];
      double4 xv[DEM_P_SWEEPW];
#pragma unroll
      for (int u = 0; u < DEM_P_SWEEPW; u++) wv[u] = (k0 + u < nown) ? ldu32_pol(P.nbr + (size_t)(k0 + u) * P.lcap + i, K.pol.first) : (unsigned)i;  // past the end: myself (never touches)
#pragma unroll
      for (int u = 0; u < DEM_P_SWEEPW; u++) xv[u] = ldg4(P.xr + (wv[u] & NBR_IDX));
      unsigned touch = 0u;
#pragma unroll
      for (int u = 0; u < DEM_P_SWEEPW; u++) {
        const double rsq = sq3_rn(xi.x - xv[u].x, xi.y - xv[u].y, xi.z - xv[u].z);
        const double radsum = xi.w + xv[u].w;
        touch |= (unsigned)((k0 + u < nown) && rsq < __dmul_rn(radsum, radsum)) << u;  // pair_gran_base.h:358
      }
#pragma unroll
      for (int u = 0; u < DEM_P_SWEEPW; u++) {
        if (!((touch >> u) & 1u)) continue;
        if (nc < DEM_P_CMAX) { s_w[nc++][tid] = wv[u]; prefetch_contact2(P, i, wv[u]); touch &= ~(1u << u); }
      }
      while (touch) {  // more than DEM_P_CMAX owned contacts (rare): evaluated on the spot
        const int u = __ffs((int)touch) - 1;
        touch &= touch - 1;
        pair_contact2<NORMAL, ROLLING, ONE>(P, K, i, P.nbr[(size_t)(k0 + u) * P.lcap + i], xi, vi, wi, su, &s_nh[tid], F, T);
      }
    }
  }
  // own-lane rounds; their number minimises (own rounds) x 3 + (cooperative rounds that remain) x 5 for this warp
  int ownr = 0;
  {
    int best = 0x7fffffff;
#pragma unroll 1
    for (int r = 0; r <= DEM_P_OWNR; r++) {
      const int rem = __reduce_add_sync(0xffffffffu, max(nc - r, 0));
      const int cost = r * DEM_COST_OWN + ((rem + 31) >> 5) * DEM_COST_COOP;
      if (cost < best) { best = cost; ownr = r; }
      if (rem == 0) break;
    }
    const double4 xo = s_rec[0][tid], vo = s_rec[1][tid], wo = s_rec[2][tid];
#pragma unroll 1
    for (int r = 0; r < ownr; r++)
      if (r < nc) pair_contact2<NORMAL, ROLLING, ONE>(P, K, i, s_w[r][tid], xo, vo, wo, su, &s_nh[tid], F, T);
  }
  {  // cooperative deal of the remaining items: item t belongs to the last lane whose first item is <= t
    const int ncc = max(nc - ownr, 0);
    int incl = ncc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int excl = incl - ncc;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    s_off[tid] = excl;
    __syncwarp();
    for (int b0 = 0; b0 < total; b0 += 32) {
      const int bend = min(total, b0 + 32);
      const int t = b0 + lane;
      if (t < total) {
        int p = 0;
#pragma unroll
        for (int s = 16; s; s >>= 1) if (s_off[wb + p + s] <= t) p += s;
        const int q = wb + p;
        const unsigned w = s_w[ownr + t - s_off[q]][q];
        double rF[3] = {0., 0., 0.}, rT[3] = {0., 0., 0.};
        pair_contact2<NORMAL, ROLLING, ONE>(P, K, i - tid + q, w, s_rec[0][q], s_rec[1][q], s_rec[2][q], su, &s_nh[q], rF, rT);
        const int sl = (wb >> 5) * 32 + (t - b0);
#pragma unroll
        for (int d = 0; d < 3; d++) { s_res[d][sl] = rF[d]; s_res[3 + d][sl] = rT[d]; }
      }
      __syncwarp();
      const int qe = min(incl, bend);
      for (int k = max(excl, b0); k < qe; k++) {
        const int sl = (wb >> 5) * 32 + (k - b0);
#pragma unroll
        for (int d = 0; d < 3; d++) { F[d] += s_res[d][sl]; T[d] += s_res[3 + d][sl]; }
      }
      __syncwarp();
    }
  }
  if (active) {
    const int nh = s_nh[tid];
    if (nh != nh0) P.numneigh[i] = NN2_PACK(NN2_TOT(nnw), nown, nh);
    double *pp = K.part + (i - K.cs);
#pragma unroll
    for (int d = 0; d < 3; d++) { st_pol(pp + (size_t)d * P.ccap, F[d], K.pol.last); st_pol(pp + (size_t)(3 + d) * P.ccap, T[d], K.pol.last); }
  }
}

// Phase B of a step for particle i: own partial sum + the shares of the pairs the partners own (row order), then the
// owner epilogue (gravity, wall forces of the pre-passes, freeze, integration, rebuild trigger)
#ifndef DEM_F_UNROLL
#define DEM_F_UNROLL 4
#endif
template <bool CG>
__device__ __forceinline__ bool finish_particle(const StepP &P, const Chunk &K, int i)
{
  const double4 xi = ldg4(P.xr + i), vi = ldg4(P.vm + i), wi = ldg4(P.wt + i);
  double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
  if (P.have_pair) {
    const int nnw = P.numneigh[i];
    const int nnon = NN2_TOT(nnw) - NN2_OWN(nnw);
    {
      const double *pp = K.part + (i - K.cs);
#pragma unroll
      for (int d = 0; d < 3; d++) { F[d] = ldcg_pol(pp + (size_t)d * P.ccap, K.pol.last); T[d] = ldcg_pol(pp + (size_t)(3 + d) * P.ccap, K.pol.last); }
    }
    for (int m0 = 0; m0 < nnon; m0 += DEM_F_UNROLL) {
      unsigned wv[DEM_F_UNROLL];
      double4 r0[DEM_F_UNROLL], r1[DEM_F_UNROLL];
#pragma unroll
      for (int u = 0; u < DEM_F_UNROLL; u++) {
        const unsigned *q = P.nbr + (size_t)(P.maxk - 1 - (m0 + u)) * P.lcap + i;
        wv[u] = (m0 + u < nnon) ? ldcgu32_pol(q, K.pol.first) : 0u;
      }
#pragma unroll
      for (int u = 0; u < DEM_F_UNROLL; u++) {
        r0[u] = make_double4(0., 0., 0., 0.); r1[u] = make_double4(0., 0., -1., 0.);
        if (wv[u] & NBR_HIST) {
          const int slot = (int)((wv[u] & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
          const int j = (int)(wv[u] & NBR_IDX);  // the owner: lower index, same or previous chunk
          const double4 *rp = j >= K.cs ? res_rec(P, K.res, slot, j - K.cs) : res_rec(P, K.res_prev, slot, j - K.cs_prev);
          r0[u] = ldcg4_pol(rp, K.pol.last); r1[u] = ldcg4_pol(rp + 1, K.pol.last);
        }
      }
#pragma unroll
      for (int u = 0; u < DEM_F_UNROLL; u++) {
        // (a ring slot is shared by the chunks c, c + ring, ...: the record must be this launch's AND this owner's)
        if ((wv[u] & NBR_HIST) && r1[u].z == P.serial && r1[u].w == (double)(wv[u] & NBR_IDX)) {
          F[0] -= r0[u].x; F[1] -= r0[u].y; F[2] -= r0[u].z;
          T[0] += r0[u].w; T[1] += r1[u].x; T[2] += r1[u].y;
        }
      }
    }
  }
  return step_epilogue(P, i, xi, vi, wi, F, T);
}

// ---- plain two-launch form (option wave 0: a single chunk; kept for per-phase profiling)
template <int NORMAL, int ROLLING, bool ONE>
__global__ void __launch_bounds__(128, DEM_P_MINBLOCKS) k_pairs(const StepP P)
{
  __shared__ unsigned s_w[DEM_P_CMAX][128];
  __shared__ double4 s_rec[3][128];
  __shared__ double s_res[6][4 * 32];
  __shared__ int s_off[128], s_nh[128];
  if (step_gated(P)) return;
  const Chunk K = chunk_of(P, 0, 0, P.nlocal, 0);
  pairs_block<NORMAL, ROLLING, ONE>(P, K, blockIdx.x * 128, s_w, s_rec, s_res, s_off, s_nh);
}
__global__ void __launch_bounds__(256) k_finish(const StepP P)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (step_gated(P)) return;
  bool trig = false;
  const Chunk K = chunk_of(P, 0, 0, P.nlocal, 0);
  if (i < P.nlocal) trig = finish_particle<false>(P, K, i);
  if (__any_sync(0xffffffffu, trig) && (threadIdx.x & 31) == 0) *((volatile int *)P.flag) = 1;
}

// ---- the step as ONE persistent kernel: blocks draw (chunk, phase, block) items from a queue whose order interleaves
// phase A of chunk t with phase B of chunk t - skew; counters of finished blocks per chunk and phase carry the
// dependencies (always on items handed out earlier, so a waiting block can never starve the block it waits for)
__device__ __forceinline__ bool wave_wait(const StepP &P, const int *ctr, int target)
{  // bounded (~0.5 s for the first block that gives up, none for the others): a broken dependency must not hang the device
  const long long t0 = clock64();
  while (*((volatile const int *)ctr) < target) {
    __nanosleep(100);
    if (((volatile int *)P.flag)[3] || clock64() - t0 > 1000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ int chunk_blocks(const StepP &P, int c) { return (P.chunk_start[c + 1] - P.chunk_start[c] + 127) >> 7; }
template <int NORMAL, int ROLLING, bool ONE>
__global__ void __launch_bounds__(128, DEM_P_MINBLOCKS) k_wave(const StepP P)
{
  __shared__ unsigned s_w[DEM_P_CMAX][128];
  __shared__ double4 s_rec[3][128];
  __shared__ double s_res[6][4 * 32];
  __shared__ int s_off[128], s_nh[128];
  __shared__ int s_item[2], s_grp[2];
  if (step_gated(P)) return;
  const int tid = threadIdx.x;
  int *ctr = P.wctr, *doneA = P.wctr + 8, *doneB = P.wctr + 8 + 256;
  // thread 0 draws the items one ahead (the atomic's latency hides behind the current item) and tracks the work group of
  // its item with a running pointer (a block's items ascend)
  int gcur = 0, nxt = 0, par = 0;
  if (tid == 0) nxt = atomicAdd(ctr, 1);
  for (;;) {
    if (tid == 0) {
      const int item = nxt;
      if (item < P.nitems) { while (P.grp_start[gcur + 1] <= item) gcur++; nxt = atomicAdd(ctr, 1); }
      s_item[par] = item; s_grp[par] = gcur;
    }
    __syncthreads();
    const int item = s_item[par], g = s_grp[par];
    par ^= 1;
    if (item >= P.nitems) break;
    const int gc = P.grp[g], c = gc >> 1, b = item - P.grp_start[g];
    const Chunk K = chunk_of(P, c, P.chunk_start[c], P.chunk_start[c + 1], c > 0 ? P.chunk_start[c - 1] : 0);
    if ((gc & 1) == 0) {
      if (c >= P.ring) {  // my ring slot was chunk c - ring's: its readers are phase B of chunks c - ring and c - ring + 1
        if (tid == 0) {
          bool ok = wave_wait(P, doneB + c - P.ring + 1, chunk_blocks(P, c - P.ring + 1));
          ok = ok && wave_wait(P, doneB + c - P.ring, chunk_blocks(P, c - P.ring));
          if (!ok) ((volatile int *)P.flag)[3] = 1;
        }
        __syncthreads();
      }
      pairs_block<NORMAL, ROLLING, ONE>(P, K, K.cs + b * 128, s_w, s_rec, s_res, s_off, s_nh);
      __syncthreads();
      if (tid == 0) { __threadfence(); atomicAdd(doneA + c, 1); }
    } else {
      if (tid == 0) {
        bool ok = wave_wait(P, doneA + c, chunk_blocks(P, c));
        if (c > 0) ok = ok && wave_wait(P, doneA + c - 1, chunk_blocks(P, c - 1));
        if (!ok) ((volatile int *)P.flag)[3] = 1;
        __threadfence();
      }
      __syncthreads();
      const int i = K.cs + b * 128 + tid;
      bool trig = false;
      if (i < K.ce) trig = finish_particle<true>(P, K, i);
      if (__any_sync(0xffffffffu, trig) && (tid & 31) == 0) *((volatile int *)P.flag) = 1;
      __syncthreads();
      if (tid == 0) { __threadfence(); atomicAdd(doneB + c, 1); }
    }
  }
  if (tid == 0) {  // the last block to leave re-arms the queue for the next launch
    __threadfence();
    if (atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1) {
      for (int k = 0; k < P.nslab; k++) { doneA[k] = 0; doneB[k] = 0; }
      ctr[1] = 0; __threadfence(); ctr[0] = 0;
    }
  }
}
// chunk table and work queue of the wavefront, from the sorted cell keys of the owned particles (rebuild time)
__global__ void __launch_bounds__(256) k_wave_table(int nslab, int n, const unsigned *skeys, const GridP G, int skew,
                                                    int *chunk_start, int *grp, int *grp_start, int *meta)
{
  __shared__ int cs[260];
  for (int c = threadIdx.x; c <= nslab; c += blockDim.x) {
    int lo = 0, hi = n;  // first particle whose key is >= the slab's first key
    if (c == nslab) lo = n;
    else { const unsigned key = slab_key_lo(G, c); while (lo < hi) { const int mid = (lo + hi) >> 1; if (skeys[mid] < key) lo = mid + 1; else hi = mid; } }
    cs[c] = lo; chunk_start[c] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int g = 0, acc = 0, mx = 0;
    grp_start[0] = 0;
    for (int t = 0; t < nslab + skew; t++) {
      if (t < nslab) { grp[g] = t * 2; acc += (cs[t + 1] - cs[t] + 127) >> 7; grp_start[++g] = acc; mx = max(mx, cs[t + 1] - cs[t]); }
      if (t >= skew) { const int c = t - skew; grp[g] = c * 2 + 1; acc += (cs[c + 1] - cs[c] + 127) >> 7; grp_start[++g] = acc; }
    }
    meta[0] = acc; meta[1] = mx; meta[2] = g;
  }
}

// ------------------------------------------------------------------------------------------------ rebuild
// Owner list + history remap: neigh_gran.cpp:560-625 (predicates), fix_contact_history.cpp:351.  One thread per particle
// walks the 27-cell stencil; owned entries grow from the front of the row, the others from the back.  A kept history row is
// looked up in the particle's OLD row by partner tag; the old word says in whose row the records are (the old owner).
__global__ void __launch_bounds__(128) k_build_list2(const BuildP B, int nlocal_old, int maxk_old)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.nlocal) return;
  const double4 xi = B.xr[i];
  const int tagi = B.tag[i];
  int cx, cy, cz; cell_of(B.G, xi, cx, cy, cz);
  const int oi = B.have_old ? B.perm[i] : 0;
  const int nnw_old = B.have_old ? B.numneigh_old[oi] : 0;
  const int nown_old = NN2_OWN(nnw_old), nnon_old = NN2_TOT(nnw_old) - nown_old;
  int nown = 0, nnon = 0, nh = 0, nband = 0;
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
            if (!(rsq <= __dmul_rn(rc, rc))) continue;
            const bool owned = pass || j > i;
            const bool inband = rsq < __dmul_rn(radsum, radsum);
            const int tagj = B.tag[j];
            unsigned w = (unsigned)j | (tagj < tagi ? NBR_JFIRST : 0u);
            if (owned && inband) {
              nband++;
              for (int m = 0; m < nown_old + nnon_old; m++) {
                const size_t eo = (size_t)(m < nown_old ? m : maxk_old - 1 - (m - nown_old)) * B.cap_old + oi;
                const unsigned wo = B.nbr_old[eo];
                if ((wo & NBR_HIST) && B.ptag_old[eo] == tagj) {
                  const int so = (int)((wo & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
                  const int jo = (int)(wo & NBR_IDX);
                  const int owner_old = (jo > oi || jo >= nlocal_old) ? oi : jo;
                  if (nh < B.hslots) {
                    w |= (unsigned)(nh + 1) << NBR_SLOT_SHIFT;
                    for (int d = 0; d < B.dnum; d++) B.hist[(size_t)(nh * B.dnum + d) * B.cap + i] = B.hist_old[(size_t)(so * B.dnum + d) * B.cap_old + owner_old];
                  }
                  nh++;
                  break;
                }
              }
            }
            if (nown + nnon < B.maxk) {
              const size_t en = (size_t)(owned ? nown : B.maxk - 1 - nnon) * B.cap + i;
              B.nbr[en] = w; B.ptag[en] = tagj;
            }
            if (owned) nown++; else nnon++;
          }
        }
      }
    }
  }
  const int tot = nown + nnon;
  B.numneigh[i] = tot <= B.maxk ? NN2_PACK(tot, nown, min(nh, B.hslots)) : 0;
  if (tot > B.maxk) atomicMax(B.overflow, tot);
  if (max(nh, nband) + 8 > B.hslots) atomicMax(B.overflow + 1, max(nh, nband));
}
// second pass of the rebuild: every entry of a pair the PARTNER owns learns the pair's history slot from the owner's row
__global__ void __launch_bounds__(128) k_link_slots(int n, int cap, int maxk, unsigned *nbr, const int *numneigh)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nnw = numneigh[i];
  const int nnon = NN2_TOT(nnw) - NN2_OWN(nnw);
  for (int m = 0; m < nnon; m++) {
    unsigned *q = nbr + (size_t)(maxk - 1 - m) * cap + i;
    const unsigned w = *q;
    const int j = (int)(w & NBR_IDX);
    const int nown_j = NN2_OWN(numneigh[j]);
    unsigned slotbits = 0u;
    for (int k = 0; k < nown_j; k++) {
      const unsigned wj = nbr[(size_t)k * cap + j];
      if ((wj & NBR_IDX) == (unsigned)i) { slotbits = wj & NBR_HIST; break; }
    }
    *q = (w & ~NBR_HIST) | slotbits;
  }
}
__device__ __forceinline__ size_t row_entry2(int m, int nown, int maxk, int cap, int i)
{  // m-th entry of a row: owned entries first, then the partner-owned ones from the back
  return (size_t)(m < nown ? m : maxk - 1 - (m - nown)) * cap + i;
}
__global__ void __launch_bounds__(256) k_count_pairs2(int n, const int *numneigh, const unsigned *nbr, int cap, int maxk, unsigned long long *out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long a = 0, b = 0;
  if (i < n) {
    const int nnw = numneigh[i], tot = NN2_TOT(nnw), nown = NN2_OWN(nnw);
    a = tot;
    for (int m = 0; m < tot; m++) b += (nbr[row_entry2(m, nown, maxk, cap, i)] & NBR_HIST) ? 1 : 0;
  }
  for (int o = 16; o; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { if (a) atomicAdd(out, a); if (b) atomicAdd(out + 1, b); }
}
// histories a migrating particle takes along: every entry with a slot, from the owner's row
__global__ void __launch_bounds__(256) k_max_nh2(int n, const int *list, const int *numneigh, const unsigned *nbr, int cap, int maxk, int *out)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int i = list[q];
  const int nnw = numneigh[i], tot = NN2_TOT(nnw), nown = NN2_OWN(nnw);
  int c = 0;
  for (int m = 0; m < tot; m++) c += (nbr[row_entry2(m, nown, maxk, cap, i)] & NBR_HIST) ? 1 : 0;
  atomicMax(out, c);
}


// migration records of the owner list (same record layout as k_mig_pack / k_mig_unpack): the leaving particle takes the
// history of EVERY pair it is part of along -- from its own row where it owns the pair, from the partner's row otherwise
// (fix_contact_history.cpp:508-555: each atom carries its partners and their history, both directions)
__global__ void __launch_bounds__(128) k_mig_pack2(const MigP M, int nlocal_old)
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
    const int nnw = M.numneigh[i], tot = NN2_TOT(nnw), nown = NN2_OWN(nnw);
    double *hb = b + 16 + M.nwrows + mblock;
    for (int m = 0; m < tot; m++) {
      const size_t en = row_entry2(m, nown, M.maxk, M.lcap, i);
      const unsigned wd = M.nbr[en];
      if (!(wd & NBR_HIST) || nh >= M.hmax) continue;
      const int slot = (int)((wd & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
      const int j = (int)(wd & NBR_IDX);
      const int owner = (j > i || j >= nlocal_old) ? i : j;
      double *e = hb + (size_t)nh * (1 + 4 * M.hrec);
      e[0] = (double)M.ptag[en];
      for (int r = 0; r < M.hrec; r++) {
        const double4 h = M.hist[(size_t)(slot * M.hrec + r) * M.lcap + owner];
        e[1 + 4 * r] = h.x; e[2 + 4 * r] = h.y; e[3 + 4 * r] = h.z; e[4 + 4 * r] = h.w;
      }
      nh++;
    }
  }
  b[15] = (double)nh;
}
// arrivals: appended behind the current particles; their histories form an all-owned row of the OLD list (partner index =
// NBR_IDX, i.e. "not a local particle of the old list") that the remap of k_build_list2 finds by partner tag
__global__ void __launch_bounds__(128) k_mig_unpack2(const MigP M, int base)
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
  if (M.mslots) M.mint[i] = 0;
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
      M.nbr[(size_t)k * M.lcap + i] = ((unsigned)(k + 1) << NBR_SLOT_SHIFT) | NBR_IDX;
      for (int r = 0; r < M.hrec; r++)
        M.hist[(size_t)(k * M.hrec + r) * M.lcap + i] = make_double4(e[1 + 4 * r], e[2 + 4 * r], e[3 + 4 * r], e[4 + 4 * r]);
    }
    M.numneigh[i] = NN2_PACK(nh, nh, nh);
  }
}

}  // namespace dem
