// dem_mesh_host.h -- host-side (setup-time) preprocessing of triangle meshes for the B200 DEM engine:
// per-triangle geometry, mesh topology (neighbours, active edges / corners, coplanar node-neighbours)
// and the coarse triangle grid.  Pure C++, included by dem_engine.cu only.
//
// What must match the reference (bit-exact bookkeeping depends on the active flags):
//   geometry        surface_mesh_I.h:302-470, multi_node_mesh_I.h:153-172
//   shared edges    multi_node_mesh_I.h:281-321 (share2Nodes), surface_mesh_I.h:1040-1144 (shareEdge, handleSharedEdge)
//   corners         surface_mesh_I.h:1148-1236 (handleCorner / checkNodeRecursive)
//   coplanarity     surface_mesh_I.h:874-958
// How it is done here (not the reference's spatial-bin pair search + recursive walks): vertices are welded
// with the reference's tolerance by a sorted sweep + union-find, shared edges come from an edge -> triangles
// map, and the fan of triangles around a vertex is a union-find component of the vertex's incidence list.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <string>
#include <vector>
#include "dem_mesh.h"

namespace dem {

struct MeshHost {
  std::string id;
  int atom_type = 1, ntri = 0, first = 0, wall = -1, moving = 0;  // moving: 0 static, 1 linear, 2 rotate
  double vel[3] = {0, 0, 0};
  double rot_origin[3] = {0, 0, 0}, rot_axis[3] = {0, 0, 1}, rot_omega = 0.;  // fix move/mesh rotate (axis normalised)
  int stress = 0; double p_ref[3] = {0, 0, 0};  // fix mesh/surface/stress: total force / torque about p_ref (mesh_module_stress.cpp)
  double curvature = 1. - 0.00001, precision = 1e-8;  // surface_mesh.h:61, multi_node_mesh.h:59
  std::vector<double> nodes;                          // [ntri][3][3] as given by the caller
  std::vector<int> edge_active, corner_active, obtuse, nneighs;  // read-back for tests (dem_download_mesh)
};

namespace meshhost {

inline void sub(const double *a, const double *b, double *r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
inline double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double mag(const double *v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline void cross(const double *a, const double *b, double *r) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; }
inline void sdiv(double *v, double s) { const double sinv = 1. / s; v[0] = sinv * v[0]; v[1] = sinv * v[1]; v[2] = sinv * v[2]; }
inline bool comp(double a, double b, double prec) { if (a == b) return true; const double x = a - b; return x < prec && x > -prec; }

inline void geometry(const double *nd, TriRec &T)
{
  for (int k = 0; k < 9; k++) T.node[k] = nd[k];
  double avg[3] = {0., 0., 0.};
  for (int i = 0; i < 3; i++) for (int d = 0; d < 3; d++) avg[d] = nd[3 * i + d] + avg[d];
  sdiv(avg, 3.0);
  double rb = 0.;
  for (int i = 0; i < 3; i++) { double v[3]; sub(avg, nd + 3 * i, v); rb = std::max(rb, mag(v)); }
  for (int d = 0; d < 3; d++) T.center[d] = avg[d];
  T.rbound = rb;
  for (int i = 0; i < 3; i++) {
    double *e = T.edgeVec + 3 * i;
    sub(nd + 3 * ((i + 1) % 3), nd + 3 * i, e);
    T.edgeLen[i] = mag(e);
    sdiv(e, T.edgeLen[i]);
  }
  cross(T.edgeVec, T.edgeVec + 3, T.surfNorm);
  sdiv(T.surfNorm, mag(T.surfNorm));
  for (int i = 0; i < 3; i++) { double *en = T.edgeNorm + 3 * i; cross(T.edgeVec + 3 * i, T.surfNorm, en); sdiv(en, mag(en)); }
}

// SurfaceMesh::recalcLocalSurfProperties (surface_mesh_I.h:187-222): what the reference recomputes from the nodes at every
// setup and, for moving meshes, at every neighbour rebuild (fix_mesh.cpp:491-575 -> pbcExchangeBorders -> refreshOwned)
inline void surf_refresh(TriRec &T)
{
  const double *nd = T.node;
  for (int i = 0; i < 3; i++) {
    double *e = T.edgeVec + 3 * i;
    sub(nd + 3 * ((i + 1) % 3), nd + 3 * i, e);
    T.edgeLen[i] = mag(e);
    sdiv(e, T.edgeLen[i]);
  }
  cross(T.edgeVec, T.edgeVec + 3, T.surfNorm);
  sdiv(T.surfNorm, mag(T.surfNorm));
  for (int i = 0; i < 3; i++) { double *en = T.edgeNorm + 3 * i; cross(T.edgeVec + 3 * i, T.surfNorm, en); sdiv(en, mag(en)); }
  int ob = -1;
  for (int i = 0; i < 3; i++) ob = dot(T.edgeVec + 3 * i, T.edgeVec + 3 * ((i + 2) % 3)) > 0. ? i : -1;
  T.flags = (T.flags & ~3) | ((ob + 1) & 3);
}

struct UF {
  std::vector<int> p;
  explicit UF(int n) : p(n) { std::iota(p.begin(), p.end(), 0); }
  int find(int a) { while (p[a] != a) { p[a] = p[p[a]]; a = p[a]; } return a; }
  void unite(int a, int b) { a = find(a); b = find(b); if (a != b) p[std::max(a, b)] = std::min(a, b); }
};

// returns "" or an error text
inline std::string derive(const double lo[3], const double hi[3], MeshHost &M, int mesh_index, std::vector<TriRec> &tri, std::vector<int> &cn)
{
  const int T = M.ntri;
  const size_t base = tri.size();
  tri.resize(base + T);
  for (int t = 0; t < T; t++) {
    TriRec &R = tri[base + t];
    geometry(&M.nodes[9 * (size_t)t], R);
    R.mesh = mesh_index; R.flags = 0;
    if (!(R.edgeLen[0] > 0 && R.edgeLen[1] > 0 && R.edgeLen[2] > 0) || !(mag(R.surfNorm) > 0.5))
      return "mesh " + M.id + ": degenerate triangle " + std::to_string(t);
  }
  TriRec *R = &tri[base];
  // 1. weld vertices (MultiNodeMesh::nodesAreEqual: every component within `precision`)
  const int NV = 3 * T;
  std::vector<int> order(NV);
  std::iota(order.begin(), order.end(), 0);
  auto vx = [&](int v) { return &M.nodes[3 * (size_t)v]; };
  std::sort(order.begin(), order.end(), [&](int a, int b) { return vx(a)[0] < vx(b)[0]; });
  UF uf(NV);
  for (int a = 0; a < NV; a++)
    for (int b = a + 1; b < NV && vx(order[b])[0] - vx(order[a])[0] < M.precision; b++) {
      const double *p = vx(order[a]), *q = vx(order[b]);
      if (comp(p[0], q[0], M.precision) && comp(p[1], q[1], M.precision) && comp(p[2], q[2], M.precision)) uf.unite(order[a], order[b]);
    }
  std::vector<int> vid(NV);
  for (int v = 0; v < NV; v++) vid[v] = uf.find(v);
  // 2. candidate pairs from the edge map, handled in (i, j<i) order
  std::map<std::pair<int, int>, std::vector<int>> edges;
  for (int t = 0; t < T; t++) for (int k = 0; k < 3; k++) {
    const int a = vid[3 * t + k], b = vid[3 * t + (k + 1) % 3];
    edges[{std::min(a, b), std::max(a, b)}].push_back(t);
  }
  std::vector<std::pair<int, int>> pairs;
  for (auto &kv : edges) { auto &l = kv.second; for (size_t a = 0; a < l.size(); a++) for (size_t b = 0; b < a; b++) if (l[a] != l[b]) pairs.push_back({std::max(l[a], l[b]), std::min(l[a], l[b])}); }
  std::sort(pairs.begin(), pairs.end());
  pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
  std::vector<int> nN(T, 0), nf((size_t)T * 5, -1);
  std::vector<unsigned char> ea((size_t)3 * T, 1), ca((size_t)3 * T, 1);
  for (auto &pr : pairs) {
    const int i = pr.first, j = pr.second;
    int i1 = -1, j1 = -1, i2 = -1, j2 = -1, ns = 0;  // share2Nodes scan order: i-nodes outer, j-nodes inner
    for (int a = 0; a < 3 && i2 < 0; a++) for (int b = 0; b < 3; b++) if (vid[3 * i + a] == vid[3 * j + b]) {
      if (ns == 0) { i1 = a; j1 = b; } else { i2 = a; j2 = b; break; }
      ns++;
    }
    if (i2 < 0) continue;
    const int iE = (2 == i1 + i2) ? 2 : std::min(i1, i2), jE = (2 == j1 + j2) ? 2 : std::min(j1, j2);
    if (nN[i] < 5) nf[(size_t)i * 5 + nN[i]] = j;
    if (nN[j] < 5) nf[(size_t)j * 5 + nN[j]] = i;
    nN[i]++; nN[j]++;
    const bool coplanar = std::fabs(dot(R[i].surfNorm, R[j].surfNorm)) >= M.curvature;
    bool overlap = false;
    if (coplanar) {  // coplanarNeighsOverlap
      double vI[3], vJ[3];
      const double *pRef = R[i].node + 3 * iE, *eN = R[i].edgeNorm + 3 * iE;
      sub(R[i].node + 3 * ((iE + 2) % 3), pRef, vI); sub(R[j].node + 3 * ((jE + 2) % 3), pRef, vJ);
      overlap = dot(vI, eN) * dot(vJ, eN) > 0.;
    }
    if (!coplanar || overlap) { ea[3 * i + iE] = 1; ea[3 * j + jE] = 0; }  // i > j: the higher id keeps the edge active (surface_mesh_I.h:1122-1135)
    else { ea[3 * i + iE] = 0; ea[3 * j + jE] = 0; }
  }
  // 3. corners: fans of triangles around each welded vertex
  std::map<int, std::vector<int>> inc;  // vertex id -> list of (3*tri + node)
  for (int v = 0; v < NV; v++) inc[vid[v]].push_back(v);
  auto listed = [&](int a, int b) { for (int k = 0; k < std::min(nN[a], 5); k++) if (nf[(size_t)a * 5 + k] == b) return true; return false; };
  for (auto &kv : inc) {
    auto &l = kv.second;
    const int n = (int)l.size();
    UF fan(n);
    for (int a = 0; a < n; a++) for (int b = 0; b < a; b++) { const int ta = l[a] / 3, tb = l[b] / 3; if (ta == tb || listed(ta, tb) || listed(tb, ta)) fan.unite(a, b); }
    for (int root = 0; root < n; root++) {
      if (fan.find(root) != root) continue;
      std::vector<int> mem;
      for (int a = 0; a < n; a++) if (fan.find(a) == root) mem.push_back(l[a]);
      bool anyActive = false, colinear = false;
      int maxId = -1;
      std::vector<const double *> ev, ep;
      for (int m : mem) {
        const int t = m / 3, k = m % 3, km = (k + 2) % 3;
        maxId = std::max(maxId, t);
        ev.push_back(R[t].edgeVec + 3 * k); ep.push_back(R[t].node + 3 * ((k + 1) % 3));
        ev.push_back(R[t].edgeVec + 3 * km); ep.push_back(R[t].node + 3 * km);
        if (ea[3 * t + k] || ea[3 * t + km]) anyActive = true;
      }
      for (size_t a = 0; a < ev.size(); a++) for (size_t b = a + 1; b < ev.size(); b++)
        if (std::fabs(dot(ev[a], ev[b])) > M.curvature &&
            !(comp(ep[a][0], ep[b][0], M.precision) && comp(ep[a][1], ep[b][1], M.precision) && comp(ep[a][2], ep[b][2], M.precision))) colinear = true;
      for (int m : mem) {
        const int t = m / 3, k = m % 3;
        const double *p = R[t].node + 3 * k;
        bool inside = true;  // Domain::is_in_subdomain with the 1e-8 border padding; outside: flag untouched (true)
        for (int d = 0; d < 3; d++) inside = inside && (p[d] >= lo[d] - 1.0e-8 && p[d] < hi[d] + 1.0e-8);
        if (!inside) continue;
        ca[3 * t + k] = (colinear || !anyActive) ? 0 : (t == maxId ? 1 : 0);
      }
    }
  }
  // 4. flags, coplanar node-neighbours (variable-length lists)
  M.edge_active.assign(3 * (size_t)T, 0); M.corner_active.assign(3 * (size_t)T, 0); M.obtuse.assign(T, -1); M.nneighs = nN;
  std::vector<std::vector<int>> cnl(T);
  for (int t = 0; t < T; t++) {
    int ob = -1;  // what survives calcObtuseAngleIndex's three overwriting calls is node 2's verdict
    for (int i = 0; i < 3; i++) ob = dot(R[t].edgeVec + 3 * i, R[t].edgeVec + 3 * ((i + 2) % 3)) > 0. ? i : -1;
    M.obtuse[t] = ob;
    int fl = (ob + 1) & 3;
    for (int k = 0; k < 3; k++) { M.edge_active[3 * t + k] = ea[3 * t + k]; M.corner_active[3 * t + k] = ca[3 * t + k]; fl |= (ea[3 * t + k] ? 1 : 0) << (2 + k); fl |= (ca[3 * t + k] ? 1 : 0) << (5 + k); }
    R[t].flags = fl;
    std::vector<int> cand;
    for (int k = 0; k < 3; k++) for (int m : inc[vid[3 * t + k]]) if (m / 3 != t) cand.push_back(m / 3);
    for (int k = 0; k < std::min(nN[t], 5); k++) cand.push_back(nf[(size_t)t * 5 + k]);
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    for (int b : cand) if (std::fabs(dot(R[t].surfNorm, R[b].surfNorm)) > M.curvature) cnl[t].push_back((int)base + b);  // ascending
  }
  // `cn` collects the lists of all meshes so far as [triangle][list]; dem_engine.cu flattens it to CSR once every mesh is derived
  for (int t = 0; t < T; t++) { cn.push_back((int)cnl[t].size()); cn.insert(cn.end(), cnl[t].begin(), cnl[t].end()); }
  return "";
}

// coarse uniform grid over the box: every triangle is listed in all cells its bounding box, grown by `margin`, overlaps
inline void build_grid(const double lo[3], const double hi[3], double cell, double margin, const std::vector<TriRec> &tri,
                       double gorg[3], double ginv[3], int gnc[3], std::vector<int> &cell_start, std::vector<int> &cell_tri)
{
  for (int pass = 0;; pass++) {
    long tot = 1;
    for (int d = 0; d < 3; d++) { int n = (int)std::floor((hi[d] - lo[d]) / cell); n = std::max(n, 1); gnc[d] = n; gorg[d] = lo[d]; ginv[d] = n / (hi[d] - lo[d]); tot *= n; }
    if (tot <= (1L << 21) || pass > 40) break;
    cell *= 1.26;
  }
  const long nc = (long)gnc[0] * gnc[1] * gnc[2];
  std::vector<int> cnt(nc + 1, 0);
  auto range = [&](const TriRec &T, int d, int &a, int &b) {
    double mn = std::min(T.node[d], std::min(T.node[3 + d], T.node[6 + d])) - margin, mx = std::max(T.node[d], std::max(T.node[3 + d], T.node[6 + d])) + margin;
    a = (int)std::floor((mn - gorg[d]) * ginv[d]) - 1; b = (int)std::floor((mx - gorg[d]) * ginv[d]) + 1;  // one guard cell against rounding
    a = std::min(std::max(a, 0), gnc[d] - 1); b = std::min(std::max(b, 0), gnc[d] - 1);
  };
  for (int pass = 0; pass < 2; pass++) {
    for (size_t t = 0; t < tri.size(); t++) {
      int a[3], b[3];
      for (int d = 0; d < 3; d++) range(tri[t], d, a[d], b[d]);
      for (int z = a[2]; z <= b[2]; z++) for (int y = a[1]; y <= b[1]; y++) for (int x = a[0]; x <= b[0]; x++) {
        const long c = ((long)z * gnc[1] + y) * gnc[0] + x;
        if (pass == 0) cnt[c + 1]++; else cell_tri[cnt[c]++] = (int)t;
      }
    }
    if (pass == 0) {
      for (long c = 0; c < nc; c++) cnt[c + 1] += cnt[c];
      cell_start.assign(cnt.begin(), cnt.end());
      cell_tri.assign(cnt[nc], 0);
    }
  }
}

}  // namespace meshhost
}  // namespace dem
