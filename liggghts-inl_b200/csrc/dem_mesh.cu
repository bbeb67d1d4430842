// dem_mesh.cu -- triangle-mesh wall kernels (sm_100a).  Compiled with --fmad=false: see dem_mesh.h.
//
// B200-first layout: the reference walks triangle-major (each triangle owns a std::vector of
// particles, fix_wall_gran.cpp:851-856) and keeps per-particle partner pages; here everything is
// particle-major so that one thread owns one particle's candidate row, contact rows and force sum:
//   rebuild : k_mesh_cand   particle -> coarse grid cell -> ascending triangle ids -> ELLPACK candidate row,
//                           plus the carry-over of the contact rows (sort_contacts, cleanUpContactJumps)
//   step    : k_mesh_step   over the compact list of wall-candidate particles (a thin layer of the bed);
//                           triangles are visited in ascending id, which is the order the reference's
//                           triangle-major loop presents them to any one particle, so the order-dependent
//                           coplanar de-duplication (fix_contact_history_mesh_I.h:61-87) is reproduced
//   motion  : k_mesh_move   nodes += v*dt, rebuild trigger on node displacement
#include "dem_mesh.h"
#include "dem_contact.cuh"

namespace dem {

#define SMALL_TRIMESH (1.e-10)
#define LARGE_TRIMESH 1000000

__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void sub3(const double *a, const double *b, double *r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }

// TriMesh::resolveTriSphereNeighbuild  tri_mesh_I.h:275-305
__device__ __forceinline__ bool tri_neighbuild(const TriRec &T, double rSphere, const double *c, double treshold)
{
  const double maxDist = rSphere + treshold;
  double v[3];
  sub3(c, T.center, v);
  if (fabs(dot3(T.surfNorm, v)) > maxDist) return false;
  const double dParaMax = maxDist * maxDist;
  for (int i = 0; i < 3; i++) {
    sub3(c, T.node + 3 * i, v);
    const double d = dot3(T.edgeNorm + 3 * i, v);
    if (d > 0 && d * d > dParaMax) return false;
  }
  return true;
}

__device__ __forceinline__ double calc_dist(const double *cs, const double *cp, double *delta)
{  // tri_mesh_I.h:309-313
  sub3(cp, cs, delta);
  return sqrt((cs[0] - cp[0]) * (cs[0] - cp[0]) + (cs[1] - cp[1]) * (cs[1] - cp[1]) + (cs[2] - cp[2]) * (cs[2] - cp[2]));
}
__device__ __forceinline__ bool edge_active(const TriRec &T, int k) { return (T.flags >> (2 + k)) & 1; }
__device__ __forceinline__ bool corner_active(const TriRec &T, int k) { return (T.flags >> (5 + k)) & 1; }

// Closest feature of a triangle for a sphere centre, as data instead of as a tree of calls.  The barycentric sign pattern of the
// centre's projection names a region of the plane (tri_mesh_I.h:65-127): 7 = over the face, 3 / 5 / 6 = beyond one edge,
// 1 / 2 / 4 = beyond a corner.  Each region resolves to ONE feature -- the face, a point at parameter s on an edge, or a node --
// through at most two projections onto edge directions (the second only at an obtuse corner, :174-255); the feature is then
// evaluated by one common tail.  Every arithmetic expression is the reference's (same operands, same order: the region codes
// and overlaps must be bit-identical); inactive edges / corners answer LARGE_TRIMESH (:141-158, skip_inactive).
struct TriFeature { int kind, idx, base; double s, sgn; };  // kind 0 node `idx`; 1 point base-node + s * edgeVec[idx], bary[base] = 1 + sgn * s / len
__device__ __forceinline__ double along(const TriRec &T, const double *p, int node, int edge)
{
  double v[3];
  sub3(p, T.node + 3 * node, v);
  return dot3(v, T.edgeVec + 3 * edge);
}
__device__ __forceinline__ TriFeature feature_beyond_edge(const TriRec &T, int e, const double *p)
{  // :129-170
  const int ip = (e + 1) % 3;
  const double s = along(T, p, e, e);
  if (s < -SMALL_TRIMESH) return TriFeature{0, e, e, 0., 0.};
  if (s > T.edgeLen[e] + SMALL_TRIMESH) return TriFeature{0, ip, ip, 0., 0.};
  return TriFeature{1, e, e, s, -1.};
}
__device__ __forceinline__ TriFeature feature_beyond_corner(const TriRec &T, int k, bool obtuse, const double *p)
{  // :174-255: at an obtuse corner the region reaches over the two adjacent edges
  const int ip = (k + 1) % 3, ipp = (k + 2) % 3;
  if (obtuse) {
    const double sb = along(T, p, k, ipp);  // backwards along the edge that ends in this node
    if (sb < SMALL_TRIMESH) return sb > -T.edgeLen[ipp] ? TriFeature{1, ipp, k, sb, 1.} : TriFeature{0, ipp, ipp, 0., 0.};
    const double sf = along(T, p, k, k);    // forwards along the edge that starts here
    if (sf > -SMALL_TRIMESH) return sf < T.edgeLen[k] ? TriFeature{1, k, k, sf, -1.} : TriFeature{0, ip, ip, 0., 0.};
  }
  return TriFeature{0, k, k, 0., 0.};
}
// TriMesh::resolveTriSphereContactBary  tri_mesh_I.h:65-127 ; returns distance - radius
__device__ double tri_contact(const TriRec &T, double precision, double rSphere, const double *c, double *delta, double *bary, int &barySign)
{
  double n0c[3];
  sub3(c, T.node, n0c);
  {  // MathExtraLiggghts::calcBaryTriCoords  math_extra_liggghts.h:583-594
    const double a = dot3(n0c, T.edgeVec), b = dot3(n0c, T.edgeVec + 6), cc = dot3(T.edgeVec, T.edgeVec + 6);
    const double oneMinCSqr = 1 - cc * cc;
    bary[1] = (a - b * cc) / (T.edgeLen[0] * oneMinCSqr);
    bary[2] = (a * cc - b) / (T.edgeLen[2] * oneMinCSqr);
    bary[0] = 1. - bary[1] - bary[2];
  }
  const double invlen = 1. / (2. * T.rbound);
  const int bs = (bary[0] > -precision * invlen) + 2 * (bary[1] > -precision * invlen) + 4 * (bary[2] > -precision * invlen);
  barySign = bs;
  if (bs == 0) return 1. - rSphere;
  if (bs == 7) {  // over the face (:259-271): foot of the perpendicular, barycentric coordinates as computed above
    const double dNorm = dot3(T.surfNorm, n0c);
    double cs[3];
    for (int k = 0; k < 3; k++) cs[k] = c[k] - T.surfNorm[k] * dNorm;
    return calc_dist(c, cs, delta) - rSphere;
  }
  // region -> (edge or corner, which): bit patterns with two set bits lie beyond the edge opposite to the cleared bit
  const int ob = (T.flags & 3) - 1;
  const int single = (bs & (bs - 1)) == 0;                       // 1, 2, 4: beyond a corner
  const int which = single ? (bs == 1 ? 0 : bs == 2 ? 1 : 2) : (bs == 3 ? 0 : bs == 6 ? 1 : 2);   // node 0 / 1 / 2 (1, 2, 4) ; edge 0 / 1 / 2 (3, 6, 5)
  const TriFeature f = single ? feature_beyond_corner(T, which, ob == which, c) : feature_beyond_edge(T, which, c);
  if (f.kind == 0) {
    if (!corner_active(T, f.idx)) return LARGE_TRIMESH - rSphere;
    bary[0] = bary[1] = bary[2] = 0.; bary[f.idx] = 1.;
    return calc_dist(c, T.node + 3 * f.idx, delta) - rSphere;
  }
  if (!edge_active(T, f.idx)) return LARGE_TRIMESH - rSphere;
  double cp[3];
  for (int d = 0; d < 3; d++) cp[d] = T.node[3 * f.base + d] + f.s * T.edgeVec[3 * f.idx + d];
  const double dist = calc_dist(c, cp, delta);
  // the two nodes of edge idx share the weight; the edge's other node is idx + 1 when the point is measured from the edge's
  // own start (sgn -1) and idx itself when it is measured backwards from the edge's end node (sgn +1, base = idx + 1)
  const int other = f.sgn < 0. ? (f.idx + 1) % 3 : f.idx;
  const int third = 3 - f.base - other;
  bary[third] = 0.;
  bary[f.base] = 1. + f.sgn * (f.s / T.edgeLen[f.idx]);
  bary[other] = 1. - bary[f.base];
  return dist - rSphere;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int *part_row(const MeshP &M, int s, int i) { return M.mint + (size_t)(1 + s) * M.cap + i; }
__device__ __forceinline__ int *cand_row(const MeshP &M, int k, int i) { return M.mint + (size_t)(1 + M.mslots + k) * M.cap + i; }
__device__ __forceinline__ double4 *hist_rec(const MeshP &M, int s, int r, int i) { return M.mhist + (size_t)(s * M.hrec + r) * M.cap + i; }
__device__ __forceinline__ void zero_hist(const MeshP &M, int s, int i)
{
  for (int r = 0; r < M.hrec; r++) *hist_rec(M, s, r, i) = make_double4(0., 0., 0., 0.);
}
__device__ __forceinline__ void swap_rows(const MeshP &M, int a, int b, int i)
{
  int *pa = part_row(M, a, i), *pb = part_row(M, b, i);
  const int t = *pa; *pa = *pb; *pb = t;
  for (int r = 0; r < M.hrec; r++) { double4 *ha = hist_rec(M, a, r, i), *hb = hist_rec(M, b, r, i); const double4 h = *ha; *ha = *hb; *hb = h; }
}

__global__ void __launch_bounds__(128) k_mesh_cand(const MeshP M, int nlocal, const double4 *xr, double skin, double cdf)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  const double4 x4 = xr[i];
  const double x[3] = {x4.x, x4.y, x4.z};
  int c[3];
  for (int d = 0; d < 3; d++) { c[d] = (int)floor((x[d] - M.gorg[d]) * M.ginv[d]); c[d] = min(max(c[d], 0), M.gnc[d] - 1); }
  const int cell = (c[2] * M.gnc[1] + c[1]) * M.gnc[0] + c[0];
  const int s = M.cell_start[cell], e = M.cell_start[cell + 1];
  int n = 0;
  for (int q = s; q < e; q++) {  // FixNeighlistMesh::checkBin fix_neighlist_mesh.cpp:313-363
    const int t = M.cell_tri[q];
    const TriRec &T = M.tri[t];
    const double thr = M.meta[T.mesh].moving ? skin : 0.5 * skin;  // :254-268
    if (tri_neighbuild(T, x4.w * cdf, x, thr)) { if (n < M.mcand) *cand_row(M, n, i) = t; n++; }
  }
  if (n > M.mcand) atomicMax(M.overflow, n);
  const int nnew = min(n, M.mcand);
  // carry the contact rows over the rebuild
  const int nold = min(M.mint[i], M.mslots);
  for (;;) {  // FixContactHistoryMesh::sort_contacts fix_contact_history_mesh.cpp:375-403
    int fe = -1, lf = -1;
    for (int j = 0; j < nold; j++) { const int p = *part_row(M, j, i); if (fe == -1 && p == -1) fe = j; if (p >= 0) lf = j; }
    if (fe > -1 && lf > -1 && fe < lf) swap_rows(M, fe, lf, i); else break;
  }
  int np = 0;
  for (int j = 0; j < nold; j++) np += *part_row(M, j, i) >= 0;
  int ip = 0;
  while (ip < np) {  // cleanUpContactJumps :467-504
    const int p = *part_row(M, ip, i);
    bool in = false;
    for (int k = 0; k < nnew; k++) in |= (*cand_row(M, k, i) == p);
    if (!in) { *part_row(M, ip, i) = -1; zero_hist(M, ip, i); swap_rows(M, ip, np - 1, i); np--; }
    else ip++;
  }
  M.mint[i] = nnew;
}

__device__ __noinline__ void mesh_chain(const StepP &P, const ModelP &m, const Contact &c, double (&h)[3], double (&g)[3], bool su, ContactOut &o)
{
  switch (m.normal * 4 + m.rolling) {
    case N_HERTZ * 4 + R_OFF: contact_chain<N_HERTZ, R_OFF, true>(P, m, c, h, g, su, o); break;
    case N_HERTZ * 4 + R_CDT: contact_chain<N_HERTZ, R_CDT, true>(P, m, c, h, g, su, o); break;
    case N_HERTZ * 4 + R_EPSD: contact_chain<N_HERTZ, R_EPSD, true>(P, m, c, h, g, su, o); break;
    case N_HERTZ * 4 + R_EPSD2: contact_chain<N_HERTZ, R_EPSD2, true>(P, m, c, h, g, su, o); break;
    case N_HOOKE * 4 + R_OFF: contact_chain<N_HOOKE, R_OFF, true>(P, m, c, h, g, su, o); break;
    case N_HOOKE * 4 + R_CDT: contact_chain<N_HOOKE, R_CDT, true>(P, m, c, h, g, su, o); break;
    case N_HOOKE * 4 + R_EPSD: contact_chain<N_HOOKE, R_EPSD, true>(P, m, c, h, g, su, o); break;
    default: contact_chain<N_HOOKE, R_EPSD2, true>(P, m, c, h, g, su, o); break;
  }
}

__device__ __forceinline__ bool coplanar_nn(const MeshP &M, int a, int b)
{  // SurfaceMesh::areCoplanarNodeNeighs surface_mesh_I.h:921-958, precomputed per triangle on the host
  int lo = M.cn[a], hi = M.cn[a + 1];  // ascending list: binary search
  while (lo < hi) { const int mid = (lo + hi) >> 1; const int q = M.cn[mid]; if (q == b) return true; if (q < b) lo = mid + 1; else hi = mid; }
  return false;
}

// FixWallGran::post_force_mesh fix_wall_gran.cpp:803-982 for the particles of the compact wall list
// The listed particles form a thin layer (46 k of the 1 M spheres of the hopper config), so the kernel is latency bound:
// MESH_G lanes share one particle (more lanes lose: every thread carries the 3 KB local frame of the triangle geometry).  Phase 1: the lanes split the particle's candidate triangles and compute the distance
// verdicts in parallel (the expensive part: tri_contact for ~10 candidates of which 0-2 are in contact); phase 2: the group's
// first lane walks the passing candidates in ascending order -- the order the reference's triangle-major loop presents them,
// which the coplanar de-duplication depends on -- and evaluates them again itself (same function, same operands, same bits).
#ifndef MESH_G
#define MESH_G 2  // measured on the hopper config (1,013,189 spheres): 1 lane 0.331, 2 lanes 0.320, 4 lanes 0.362, 8 lanes 0.457 ms per step (0.382 before the two-phase form)
#endif
__global__ void __launch_bounds__(128) k_mesh_step(const StepP P, const MeshP M)
{
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int cidx = gid / MESH_G, g = gid % MESH_G;
  const bool gated = P.gate && (((P.gate_mask & 1) && P.gate[0]) || ((P.gate_mask & 4) && P.gate[2]));  // speculative launch, see StepP::gate
  const int i = (cidx < P.nwc && !gated) ? P.wlist[cidx] : -1;
  const int nct = i >= 0 ? M.mint[i] : 0;
  double4 xi = make_double4(0., 0., 0., 1.);
  if (nct) xi = P.xr[i];
  const double pos[3] = {xi.x, xi.y, xi.z};
  const double radi = xi.w;
  unsigned long long pass = 0ull;
  for (int k = g; k < min(nct, 64); k += MESH_G) {
    const int t = *cand_row(M, k, i);
    const TriRec &T = M.tri[t];
    double delta[3], bary[3];
    int barysign = -1;
    const double deltan = tri_contact(T, M.meta[T.mesh].precision, radi, pos, delta, bary, barysign);
    if (!(deltan > P.cutneighmax) && (deltan <= 0 || deltan < (P.cdf - 1.0) * radi)) pass |= 1ull << k;
  }
#pragma unroll
  for (int o = 1; o < MESH_G; o <<= 1) pass |= __shfl_xor_sync(0xffffffffu, pass, o);
  if (!nct || g) return;
  const double4 vi = P.vm[i], wi = P.wt[i];
  const int itype = (int)(__double_as_longlong(wi.w) & 0xff);
  const bool su = (P.mode != MODE_SETUP);
  const int nrows = min(nct, M.mslots);  // the reference sizes the rows by the candidate count
  unsigned keep = 0;                     // markAllContacts: nothing kept yet
  double F[3] = {0., 0., 0.}, Tq[3] = {0., 0., 0.};
  for (int k = 0; k < nct; k++) {
    if (k < 64) {  // jump to the next candidate that passed
      const unsigned long long rest = pass >> k;
      if (!rest) { if (nct <= 64) break; k = 63; continue; }
      k += __ffsll((long long)rest) - 1;
    }
    const int t = *cand_row(M, k, i);
    const TriRec &T = M.tri[t];
    const MeshMeta &mm = M.meta[T.mesh];
    double delta[3], bary[3];
    int barysign = -1;
    const double deltan = tri_contact(T, mm.precision, radi, pos, delta, bary, barysign);
    if (deltan > P.cutneighmax) continue;
    const bool intersect = (deltan <= 0);
    if (!(deltan <= 0 || deltan < (P.cdf - 1.0) * radi)) continue;
    // FixContactHistoryMesh::handleContact fix_contact_history_mesh_I.h:51-87
    int slot = -1;
    for (int s = 0; s < nrows; s++) if (*part_row(M, s, i) == t) { slot = s; break; }
    if (slot < 0) {
      const bool faceflag = (7 == barysign);
      if (faceflag) {  // coplanarContactAlready :118-137
        bool already = false;
        for (int s = 0; s < nrows; s++) { const int q = *part_row(M, s, i); if (q >= 0 && q != t && ((keep >> s) & 1u) && coplanar_nn(M, q, t)) { already = true; break; } }
        if (already) continue;
      }
      for (int s = 0; s < nrows; s++) if (*part_row(M, s, i) == -1) { slot = s; break; }  // addNewTriContactToExistingParticle :166-208
      if (slot < 0) { atomicMax(M.overflow + 1, nct + 1); continue; }
      *part_row(M, slot, i) = t;
      zero_hist(M, slot, i);
      if (faceflag)  // checkCoplanarContactHistory :141-160 (the last coplanar partner wins)
        for (int s = 0; s < nrows; s++) { const int q = *part_row(M, s, i); if (q >= 0 && q != t && coplanar_nn(M, q, t)) for (int r = 0; r < M.hrec; r++) *hist_rec(M, slot, r, i) = *hist_rec(M, s, r, i); }
    }
    keep |= 1u << slot;
    const ModelP &wm = M.wm[T.mesh];
    if (intersect) {  // Walls::Granular::compute_force fix_wall_gran_base.h:159-367
      Contact c;
      c.dx = -delta[0]; c.dy = -delta[1]; c.dz = -delta[2];
      c.radi = radi; c.radj = 0.0; c.radsum = radi; c.deltan_in = -deltan;
      c.r = c.radi - c.deltan_in; c.rinv = 1.0 / c.r;
      c.meff = vi.w; c.mi = vi.w; c.mj = 0.0;
      c.vi[0] = vi.x; c.vi[1] = vi.y; c.vi[2] = vi.z;
      c.wi[0] = wi.x; c.wi[1] = wi.y; c.wi[2] = wi.z;
      c.wj[0] = c.wj[1] = c.wj[2] = 0.0;
      for (int d = 0; d < 3; d++) c.vj[d] = 0.0;
      if (mm.moving == 1 && su)  // per-node mesh velocity is zero during setup (fix_move_mesh.cpp:194-217), v_node = 0 + vel afterwards
        for (int d = 0; d < 3; d++) { const double vn = 0. + mm.vel[d]; c.vj[d] = (bary[0] * vn + bary[1] * vn + bary[2] * vn); }
      else if (mm.moving == 2 && su) {  // v_node = 0 + omegaVec x (node - reference point), mesh_mover_rotation.cpp:113-124
        double vn[3][3];
        for (int j = 0; j < 3; j++) {
          const double rx = T.node[3 * j] - mm.rot_origin[0], ry = T.node[3 * j + 1] - mm.rot_origin[1], rz = T.node[3 * j + 2] - mm.rot_origin[2];
          vn[j][0] = 0. + (mm.rot_omegavec[1] * rz - mm.rot_omegavec[2] * ry);
          vn[j][1] = 0. + (mm.rot_omegavec[2] * rx - mm.rot_omegavec[0] * rz);
          vn[j][2] = 0. + (mm.rot_omegavec[0] * ry - mm.rot_omegavec[1] * rx);
        }
        for (int d = 0; d < 3; d++) c.vj[d] = (bary[0] * vn[0][d] + bary[1] * vn[1][d] + bary[2] * vn[2][d]);
      }
      c.itype = itype; c.jtype = mm.atom_type;
      double h[3] = {0., 0., 0.}, g[3] = {0., 0., 0.};
      if (wm.rec_shear >= 0) { const double4 v = *hist_rec(M, slot, wm.rec_shear, i); h[0] = v.x; h[1] = v.y; h[2] = v.z; }
      if (wm.rec_roll >= 0) { const double4 v = *hist_rec(M, slot, wm.rec_roll, i); g[0] = v.x; g[1] = v.y; g[2] = v.z; }
      ContactOut o;
      mesh_chain(P, wm, c, h, g, su, o);
      F[0] += o.F[0]; F[1] += o.F[1]; F[2] += o.F[2];
      if (mm.stress && M.mforce) {  // MeshModuleStress::add_particle_contribution mesh_module_stress.cpp:316-345 (an output: summation order is free)
        const double frc[3] = {-o.F[0], -o.F[1], -o.F[2]};
        const double *pr = M.mpref + 3 * T.mesh;
        const double a0 = (pos[0] + delta[0]) - pr[0], a1 = (pos[1] + delta[1]) - pr[1], a2 = (pos[2] + delta[2]) - pr[2];  // contact point - p_ref
        double *acc = M.mforce + 6 * T.mesh;
        atomicAdd(acc + 0, frc[0]); atomicAdd(acc + 1, frc[1]); atomicAdd(acc + 2, frc[2]);
        atomicAdd(acc + 3, a1 * frc[2] - a2 * frc[1]); atomicAdd(acc + 4, a2 * frc[0] - a0 * frc[2]); atomicAdd(acc + 5, a0 * frc[1] - a1 * frc[0]);
      }
      Tq[0] += o.Ti[0]; Tq[1] += o.Ti[1]; Tq[2] += o.Ti[2];
      if (su) {
        if (wm.rec_shear >= 0) *hist_rec(M, slot, wm.rec_shear, i) = make_double4(h[0], h[1], h[2], 0.);
        if (wm.rec_roll >= 0) *hist_rec(M, slot, wm.rec_roll, i) = make_double4(g[0], g[1], g[2], 0.);
      }
    } else zero_hist(M, slot, i);  // surfacesClose: tangential / rolling history zeroed
  }
  for (int s = 0; s < nrows; s++)  // cleanUpContacts fix_contact_history_mesh.cpp:437-463
    if (!((keep >> s) & 1u) && *part_row(M, s, i) != -1) { *part_row(M, s, i) = -1; zero_hist(M, s, i); }
#pragma unroll
  for (int d = 0; d < 3; d++) { P.fw[(size_t)d * P.nwcap + cidx] += F[d]; P.fw[(size_t)(3 + d) * P.nwcap + cidx] += Tq[d]; }
}

// MathExtra::quatquat (math_extra.h:596-602) and MathExtraLiggghts::vec_quat_rotate (math_extra_liggghts.h:457-474),
// same expression order (this translation unit is compiled without FMA contraction)
__device__ __forceinline__ void quatquat(const double *a, const double *b, double *c)
{
  c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  c[1] = a[0] * b[1] + b[0] * a[1] + a[2] * b[3] - a[3] * b[2];
  c[2] = a[0] * b[2] + b[0] * a[2] + a[3] * b[1] - a[1] * b[3];
  c[3] = a[0] * b[3] + b[0] * a[3] + a[1] * b[2] - a[2] * b[1];
}
__device__ __forceinline__ void vec_quat_rotate(double *vec, const double *quat)
{
  const double vecQ[4] = {0., vec[0], vec[1], vec[2]}, quatC[4] = {quat[0], -quat[1], -quat[2], -quat[3]};
  double temp[4], resultQ[4];
  quatquat(quat, vecQ, temp); quatquat(temp, quatC, resultQ);
  vec[0] = resultQ[1]; vec[1] = resultQ[2]; vec[2] = resultQ[3];
}
// fix move/mesh linear: node += vel*dt, center += vel*dt ; rotate: nodes turned about the axis through the origin by the
// per-step quaternion, center re-averaged, edge vectors / edge normals / surface normal turned by the same quaternion ;
// a node that moved more than skin/2 since the last rebuild raises the flag (MultiNodeMesh::decideRebuild)
__global__ void __launch_bounds__(128) k_mesh_move(const MeshP M, int mesh, double dt, double trigsq, int *flag, const int *gate, int gate_mask)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const MeshMeta &mm = M.meta[mesh];
  if (q >= mm.ntri) return;
  if (gate && (((gate_mask & 1) && gate[0]) || ((gate_mask & 4) && gate[2]))) return;
  const int t = mm.first + q;
  TriRec &T = M.tri[t];
  bool trig = false;
  if (q == 0 && mm.stress && M.mpref) {  // the reference point is a mesh property: it travels with the mesh
    double *pr = M.mpref + 3 * mesh;
    if (mm.moving == 2) {
      double v[3] = {pr[0], pr[1], pr[2]};
      if (mm.rot_trans) for (int d = 0; d < 3; d++) v[d] = v[d] + (-mm.rot_origin[d]);
      vec_quat_rotate(v, mm.rot_dq);
      if (mm.rot_trans) for (int d = 0; d < 3; d++) v[d] = v[d] + mm.rot_origin[d];
      pr[0] = v[0]; pr[1] = v[1]; pr[2] = v[2];
    } else for (int d = 0; d < 3; d++) pr[d] = pr[d] + mm.vel[d] * dt;
  }
  if (mm.moving == 2) {
    double c[3] = {0., 0., 0.};
    for (int j = 0; j < 3; j++) {
      double nd[3] = {T.node[3 * j], T.node[3 * j + 1], T.node[3 * j + 2]};
      if (mm.rot_trans) for (int d = 0; d < 3; d++) nd[d] = nd[d] - mm.rot_origin[d];
      vec_quat_rotate(nd, mm.rot_dq);
      if (mm.rot_trans) for (int d = 0; d < 3; d++) nd[d] = nd[d] + mm.rot_origin[d];
      double dd[3];
      for (int d = 0; d < 3; d++) { T.node[3 * j + d] = nd[d]; c[d] = nd[d] + c[d]; dd[d] = nd[d] - M.nodes_last[(size_t)t * 9 + 3 * j + d]; }
      if (dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2] > trigsq) trig = true;
    }
    const double sinv = 1. / 3.;
    for (int d = 0; d < 3; d++) T.center[d] = sinv * c[d];
    for (int j = 0; j < 3; j++) { vec_quat_rotate(T.edgeVec + 3 * j, mm.rot_dq); vec_quat_rotate(T.edgeNorm + 3 * j, mm.rot_dq); }
    vec_quat_rotate(T.surfNorm, mm.rot_dq);
    if (trig) ((volatile int *)flag)[2] = 1;
    return;
  }
  double dx[3];
  for (int d = 0; d < 3; d++) dx[d] = mm.vel[d] * dt;
  for (int j = 0; j < 3; j++) {
    double dd[3];
    for (int d = 0; d < 3; d++) { T.node[3 * j + d] = T.node[3 * j + d] + dx[d]; dd[d] = T.node[3 * j + d] - M.nodes_last[(size_t)t * 9 + 3 * j + d]; }
    if (dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2] > trigsq) trig = true;
  }
  for (int d = 0; d < 3; d++) T.center[d] = T.center[d] + dx[d];
  if (trig) ((volatile int *)flag)[2] = 1;
}
__global__ void __launch_bounds__(128) k_mesh_hold(const MeshP M)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.ntri) return;
  for (int k = 0; k < 9; k++) M.nodes_last[(size_t)t * 9 + k] = M.tri[t].node[k];
}

void mesh_launch_candidates(const MeshP &M, int nlocal, const double4 *xr, double skin, double cdf, cudaStream_t st)
{
  if (nlocal > 0) k_mesh_cand<<<(nlocal + 127) / 128, 128, 0, st>>>(M, nlocal, xr, skin, cdf);
}
void mesh_launch_step(const StepP &P, const MeshP &M, cudaStream_t st)
{
  if (P.nwc > 0) k_mesh_step<<<(unsigned)(((size_t)P.nwc * MESH_G + 127) / 128), 128, 0, st>>>(P, M);
}
void mesh_launch_move(const MeshP &M, int mesh, double dt, double trigsq, int *flag, const int *gate, int gate_mask, cudaStream_t st)
{
  const int n = M.meta[mesh].ntri;
  if (n > 0) k_mesh_move<<<(n + 127) / 128, 128, 0, st>>>(M, mesh, dt, trigsq, flag, gate, gate_mask);
}
void mesh_launch_hold(const MeshP &M, cudaStream_t st)
{
  if (M.ntri > 0) k_mesh_hold<<<(M.ntri + 127) / 128, 128, 0, st>>>(M);
}

}  // namespace dem
