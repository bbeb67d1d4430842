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

}
