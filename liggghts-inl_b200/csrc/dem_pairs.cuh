// dem_pairs.cuh -- the pair step of the plain contact models as an OWNER LIST (option owner_list): a measured alternative
// to the full list of k_step, kept because it is what the reference's half list maps to and because its rebuild is cheaper.
//
// Every pair has ONE owner -- the particle with the lower storage index, or the local particle when the partner is a ghost
// -- which evaluates the contact once (k_pairs), keeps the single copy of its history and leaves the partner's share (the
// force on the owner, the torque on the partner) in a per-contact result record; the partner picks the record up through
// its own row in k_finish, which also integrates.  Nothing is accumulated with atomics: a particle's sum always runs over
// its row in row order (own contacts first, then the received shares), so runs stay bit reproducible -- what the reference
// gets from its half list with `newton off` (pair_gran_base.h:257-496, f[j] updated by the owner of the pair) without races.
//
// Measured on the 4,194,304-sphere bed (profiles/r2*_owner_list*): k_pairs 0.78 ms + k_finish 0.43 ms = 1.22-1.24 ms per step
// against 1.07 ms for the full list.  Half the contact math (364 M instead of 561 M warp instructions) does not pay for the
// result records: 64 B written and read per contact is the traffic the second history copy of the full list cost, plus the
// partial sums' round trip -- 4.2 GB per step against 2.8 GB.  A persistent wavefront form (phase B of a slab of the bed
// following its phase A through a queue, result rings meant to stay in L2, eviction hints) was built and measured at
// 1.7-2.2 ms with unchanged DRAM traffic: the ~95 k particles in flight move ~100 MB between a record's write and its read,
// the size of the L2 (commit e91a57a has that kernel).  The full list remains the default.
//
// Row layout (ELLPACK, transposed, stride lcap; numneigh word = NN2_PACK(total, owned, history slots in use)):
//   entries [0, owned)              pairs this particle owns: partner index > own index, or partner is a ghost
//   entries [maxk - (total-owned), maxk)  pairs the partner owns, filled from the back of the row
// A neighbour word is [31] partner tag < own tag, [30:25] history slot + 1 of the pair IN THE OWNER'S ROW (0: the pair has
// no history == reference contact_flag 0), [24:0] partner index.  History records live at hist[(slot*hrec + r)*lcap + owner],
// result records at res[(slot*lcap + owner)*2 + h] (one aligned 64-byte block per pair): h=0 (Fx, Fy, Fz, Tpx), h=1 (Tpy,
// Tpz, serial, 0) with F the force on the owner and Tp the torque on the partner; `serial` = the launch that wrote the record
// (a partner ignores stale records).
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

__device__ __forceinline__ double4 *res_rec(const StepP &P, int slot, int owner) { return P.res + ((size_t)slot * P.lcap + owner) * 2; }

// one owned, touching pair of particle i: evaluated once, in the owner's orientation (pair_chain); history in the canonical
// orientation "lower tag first"; the partner's share goes to the pair's result record
#ifndef DEM_P_INLINE
#define DEM_P_INLINE __forceinline__
#endif
template <int NORMAL, int ROLLING, bool ONE>
__device__ DEM_P_INLINE void pair_contact2(const StepP &P, int i, unsigned w, const double4 &xi, const double4 &vi, const double4 &wi,
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
    if (P.pm.tangential) hs = hp[(size_t)P.pm.rec_shear * P.lcap];
    if (HAS_ROLL_HIST) hr = hp[(size_t)P.pm.rec_roll * P.lcap];
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
      if (P.pm.tangential) st4(hp + (size_t)P.pm.rec_shear * P.lcap, make_double4(sgn * h[0], sgn * h[1], sgn * h[2], 0.));
      if (HAS_ROLL_HIST) st4(hp + (size_t)P.pm.rec_roll * P.lcap, make_double4(sgn * g[0], sgn * g[1], sgn * g[2], 0.));
    }
    if (j < P.nlocal) {
      double4 *rp = res_rec(P, slot, i);
      st4(rp, make_double4(Fc[0], Fc[1], Fc[2], Tp[0]));
      st4(rp + 1, make_double4(Tp[1], Tp[2], P.serial, 0.));
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
__device__ __forceinline__ void pairs_block(const StepP &P, int i0, unsigned (*s_w)[128], double4 (*s_rec)[128], double (*s_res)[4 * 32], int *s_off, int *s_nh)
{
  const int tid = threadIdx.x, lane = tid & 31, wb = tid & ~31;
  const int i = i0 + tid;
  const bool active = i < P.nlocal;
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
      for (int u = 0; u < DEM_P_SWEEPW; u++) wv[u] = (k0 + u < nown) ? P.nbr[(size_t)(k0 + u) * P.lcap + i] : (unsigned)i;  // past the end: myself (never touches)
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
        pair_contact2<NORMAL, ROLLING, ONE>(P, i, P.nbr[(size_t)(k0 + u) * P.lcap + i], xi, vi, wi, su, &s_nh[tid], F, T);
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
      if (r < nc) pair_contact2<NORMAL, ROLLING, ONE>(P, i, s_w[r][tid], xo, vo, wo, su, &s_nh[tid], F, T);
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
        pair_contact2<NORMAL, ROLLING, ONE>(P, i - tid + q, w, s_rec[0][q], s_rec[1][q], s_rec[2][q], su, &s_nh[q], rF, rT);
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
    P.f[i] = F[0]; P.f[P.cap + i] = F[1]; P.f[2 * (size_t)P.cap + i] = F[2];
    P.tq[i] = T[0]; P.tq[P.cap + i] = T[1]; P.tq[2 * (size_t)P.cap + i] = T[2];
  }
}

// Phase B of a step for particle i: own partial sum + the shares of the pairs the partners own (row order), then the
// owner epilogue (gravity, wall forces of the pre-passes, freeze, integration, rebuild trigger)
#ifndef DEM_F_UNROLL
#define DEM_F_UNROLL 4
#endif
__device__ __forceinline__ bool finish_particle(const StepP &P, int i)
{
  const double4 xi = ldg4(P.xr + i), vi = ldg4(P.vm + i), wi = ldg4(P.wt + i);
  double F[3] = {0., 0., 0.}, T[3] = {0., 0., 0.};
  if (P.have_pair) {
    const int nnw = P.numneigh[i];
    const int nnon = NN2_TOT(nnw) - NN2_OWN(nnw);
    F[0] = P.f[i]; F[1] = P.f[P.cap + i]; F[2] = P.f[2 * (size_t)P.cap + i];
    T[0] = P.tq[i]; T[1] = P.tq[P.cap + i]; T[2] = P.tq[2 * (size_t)P.cap + i];
    for (int m0 = 0; m0 < nnon; m0 += DEM_F_UNROLL) {
      unsigned wv[DEM_F_UNROLL];
      double4 r0[DEM_F_UNROLL], r1[DEM_F_UNROLL];
#pragma unroll
      for (int u = 0; u < DEM_F_UNROLL; u++) {
        const unsigned *q = P.nbr + (size_t)(P.maxk - 1 - (m0 + u)) * P.lcap + i;
        wv[u] = (m0 + u < nnon) ? *q : 0u;
      }
#pragma unroll
      for (int u = 0; u < DEM_F_UNROLL; u++) {
        r0[u] = make_double4(0., 0., 0., 0.); r1[u] = make_double4(0., 0., -1., 0.);
        if (wv[u] & NBR_HIST) {
          const int slot = (int)((wv[u] & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
          const double4 *rp = res_rec(P, slot, (int)(wv[u] & NBR_IDX));
          r0[u] = rp[0]; r1[u] = rp[1];
        }
      }
#pragma unroll
      for (int u = 0; u < DEM_F_UNROLL; u++) {
        if ((wv[u] & NBR_HIST) && r1[u].z == P.serial) {
          F[0] -= r0[u].x; F[1] -= r0[u].y; F[2] -= r0[u].z;
          T[0] += r0[u].w; T[1] += r1[u].x; T[2] += r1[u].y;
        }
      }
    }
  }
  return step_epilogue(P, i, xi, vi, wi, F, T);
}

// ---- the step as two launches: phase A over all particles, then phase B
template <int NORMAL, int ROLLING, bool ONE>
__global__ void __launch_bounds__(128, DEM_P_MINBLOCKS) k_pairs(const StepP P)
{
  __shared__ unsigned s_w[DEM_P_CMAX][128];
  __shared__ double4 s_rec[3][128];
  __shared__ double s_res[6][4 * 32];
  __shared__ int s_off[128], s_nh[128];
  if (step_gated(P)) return;
  pairs_block<NORMAL, ROLLING, ONE>(P, blockIdx.x * 128, s_w, s_rec, s_res, s_off, s_nh);
}
__global__ void __launch_bounds__(256) k_finish(const StepP P)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (step_gated(P)) return;
  bool trig = false;
  if (i < P.nlocal) trig = finish_particle(P, i);
  if (__any_sync(0xffffffffu, trig) && (threadIdx.x & 31) == 0) *((volatile int *)P.flag) = 1;
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


// ---- per-contact output (the reference: compute pair/gran/local, compute_pair_gran_local.cpp:287-303 and
// post_force_pp :519-640 -- one row per touching pair with the ids and the force / torque the pair loop applied to the
// first particle).  With option contact_output the step kernel leaves, in the force evaluations that materialise forces
// (Verlet::setup and the last step of a run), the force and torque of every evaluated contact in the pair's history slot of
// a side array, stamped with the launch serial (k_step: pair_contact).  These two kernels gather the stamped records into
// rows (own tag, partner tag, force on me, torque on me) -- both directions of a pair between two owned particles, like the
// two particles' views of it.
__global__ void __launch_bounds__(256) k_contact_count(const StepP P, int *cnt)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nlocal) return;
  const int nn = P.numneigh[i] & 0xffff;
  int c = 0;
  for (int k = 0; k < nn; k++) {
    const unsigned w = P.nbr[(size_t)k * P.lcap + i];
    if (!(w & NBR_HIST)) continue;
    const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
    c += P.cout[((size_t)slot * P.lcap + i) * 2 + 1].z == P.serial ? 1 : 0;
  }
  cnt[i] = c;
}
__global__ void __launch_bounds__(256) k_contact_fill(const StepP P, const int *ptag, const int *tag, const int *off, int *out_tags, double *out6)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nlocal) return;
  const int nn = P.numneigh[i] & 0xffff;
  int row = off[i];
  for (int k = 0; k < nn; k++) {
    const unsigned w = P.nbr[(size_t)k * P.lcap + i];
    if (!(w & NBR_HIST)) continue;
    const int slot = (int)((w & NBR_HIST) >> NBR_SLOT_SHIFT) - 1;
    const double4 a = P.cout[((size_t)slot * P.lcap + i) * 2], b = P.cout[((size_t)slot * P.lcap + i) * 2 + 1];
    if (b.z != P.serial) continue;
    out_tags[2 * row] = tag[i]; out_tags[2 * row + 1] = ptag[(size_t)k * P.lcap + i];
    double *o = out6 + 6 * (size_t)row;
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y;
    row++;
  }
}

}  // namespace dem
