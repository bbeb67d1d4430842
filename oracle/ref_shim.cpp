// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE (not product code).
// Extra extern "C" accessors linked into oracle/_ref/libliggghts_ref.so next to the
// reference's own C API (library.h). They only READ public members of the reference's
// objects so that parity tests can compare neighbour pair sets and contact-history
// bookkeeping (gate (i)):
//   Pair::list / Pair::listgranhistory            pair.h:120-123
//   NeighList::{inum,ilist,numneigh,firstneigh,firstdouble,dnum}  neigh_list.h
#include "lammps.h"
#include "atom.h"
#include "force.h"
#include "pair.h"
#include "neighbor.h"
#include "neigh_list.h"
#include "update.h"
#include "lmptype.h"
#include "modify.h"
#include "fix.h"
#include "fix_mesh_surface.h"
#include "fix_contact_history_mesh.h"
#include "tri_mesh.h"
#include <string.h>

using namespace LAMMPS_NS;

extern "C" {

// number of (i,j) entries in the granular half list, and dnum of the history list
int ref_pairlist_count(void *ptr, int *dnum)
{
  LAMMPS *lmp = (LAMMPS *) ptr;
  if (!lmp->force->pair || !lmp->force->pair->list) return -1;
  NeighList *list = lmp->force->pair->list;
  NeighList *hl = lmp->force->pair->listgranhistory;
  if (dnum) *dnum = hl ? hl->dnum : 0;
  int n = 0;
  for (int ii = 0; ii < list->inum; ii++) n += list->numneigh[list->ilist[ii]];
  return n;
}

// fill tag_i,tag_j,(local index i,j),flag and hist[dnum] per entry, in list order
int ref_pairlist_dump(void *ptr, int *tag_i, int *tag_j, int *idx_i, int *idx_j, int *flag, double *hist)
{
  LAMMPS *lmp = (LAMMPS *) ptr;
  NeighList *list = lmp->force->pair->list;
  NeighList *hl = lmp->force->pair->listgranhistory;
  const int dnum = hl ? hl->dnum : 0;
  int *tag = lmp->atom->tag;
  int n = 0;
  for (int ii = 0; ii < list->inum; ii++) {
    const int i = list->ilist[ii];
    const int *jlist = list->firstneigh[i];
    const int jnum = list->numneigh[i];
    for (int jj = 0; jj < jnum; jj++) {
      const int j = jlist[jj] & NEIGHMASK;
      tag_i[n] = tag[i]; tag_j[n] = tag[j];
      if (idx_i) idx_i[n] = i;
      if (idx_j) idx_j[n] = j;
      if (flag) flag[n] = hl ? hl->firstneigh[i][jj] : 0;
      if (hist && hl) for (int d = 0; d < dnum; d++) hist[(size_t)n*dnum + d] = hl->firstdouble[i][dnum*jj + d];
      n++;
    }
  }
  return n;
}

int ref_nlocal(void *ptr) { return ((LAMMPS *) ptr)->atom->nlocal; }
int ref_nghost(void *ptr) { return ((LAMMPS *) ptr)->atom->nghost; }
int ref_neigh_ncalls(void *ptr) { return ((LAMMPS *) ptr)->neighbor->ncalls; }
long ref_ntimestep(void *ptr) { return (long) ((LAMMPS *) ptr)->update->ntimestep; }


// ---- triangle meshes (fix mesh/surface): topology flags and per-particle mesh contact rows
//   SurfaceMesh::edgeActive/cornerActive/nNeighs (surface_mesh.h:152-162), MultiNodeMesh::node (multi_node_mesh.h:151),
//   FixContactHistory::n_partner/partner/contacthistory (fix_contact_history.h:96-110), FixContactHistoryMesh::nneighs
static FixMeshSurface *ref_find_mesh(void *ptr, const char *id)
{
  LAMMPS *lmp = (LAMMPS *) ptr;
  Fix *f = lmp->modify->find_fix_id(id);
  return f ? dynamic_cast<FixMeshSurface *>(f) : NULL;
}
int ref_mesh_ntri(void *ptr, const char *id)
{
  FixMeshSurface *fm = ref_find_mesh(ptr, id);
  return fm ? fm->triMesh()->sizeLocal() : -1;
}
// derived per-triangle geometry (surface_mesh.h:143-147,259-262) and the per-node mesh velocity "v" of fix move/mesh
int ref_mesh_geometry(void *ptr, const char *id, double *edgeVec, double *edgeNorm, double *surfNorm, double *center, double *vnode)
{
  FixMeshSurface *fm = ref_find_mesh(ptr, id);
  if (!fm) return -1;
  TriMesh *m = fm->triMesh();
  const int n = m->sizeLocal();
  MultiVectorContainer<double,3,3> *en = m->prop().getElementProperty<MultiVectorContainer<double,3,3> >("edgeNorm");
  MultiVectorContainer<double,3,3> *v = m->prop().getElementProperty<MultiVectorContainer<double,3,3> >("v");
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < 3; j++) {
      m->edgeVec(i, j, edgeVec + 9 * (size_t)i + 3 * j);
      for (int d = 0; d < 3; d++) {
        edgeNorm[9 * (size_t)i + 3 * j + d] = en ? (*en)(i)[j][d] : 0.0;
        vnode[9 * (size_t)i + 3 * j + d] = v ? (*v)(i)[j][d] : 0.0;
      }
    }
    m->surfaceNorm(i, surfNorm + 3 * (size_t)i);
    m->center(i, center + 3 * (size_t)i);
  }
  return n;
}
int ref_mesh_topology(void *ptr, const char *id, double *nodes, int *edgeActive, int *cornerActive, int *nneighs)
{
  FixMeshSurface *fm = ref_find_mesh(ptr, id);
  if (!fm) return -1;
  TriMesh *m = fm->triMesh();
  const int n = m->sizeLocal();
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < 3; j++) {
      m->node(i, j, nodes + 9 * (size_t)i + 3 * j);
      edgeActive[3 * i + j] = m->edgeActive(i, j) ? 1 : 0;
      cornerActive[3 * i + j] = m->cornerActive(i, j) ? 1 : 0;
    }
    nneighs[i] = m->nNeighs(i);
  }
  return n;
}
int ref_mesh_contacts(void *ptr, const char *id, int maxrows, int *tag, int *tri, double *hist, int *dnum)
{
  LAMMPS *lmp = (LAMMPS *) ptr;
  FixMeshSurface *fm = ref_find_mesh(ptr, id);
  if (!fm || !fm->contactHistory()) return -1;
  FixContactHistoryMesh *ch = fm->contactHistory();
  const int dn = ch->get_dnum();
  if (dnum) *dnum = dn;
  int n = 0;
  for (int i = 0; i < lmp->atom->nlocal; i++) {
    const int nn = ch->nneighs(i);
    for (int j = 0; j < nn; j++) {
      const int t = ch->partner(i, j);
      if (t < 0) continue;
      if (n < maxrows) {
        tag[n] = lmp->atom->tag[i]; tri[n] = t;
        for (int d = 0; d < dn; d++) hist[(size_t)n * dn + d] = ch->contacthistory(i, j)[d];
      }
      n++;
    }
  }
  return n;
}

}
