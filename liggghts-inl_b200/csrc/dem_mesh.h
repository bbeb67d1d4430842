// dem_mesh.h -- triangle-mesh walls of the B200 DEM engine: data shared between the host
// orchestration (dem_engine.cu) and the mesh kernels (dem_mesh.cu, compiled with --fmad=false so
// that every geometric predicate -- region codes, active-feature branches, touch tests -- rounds
// exactly like the reference's C++ and the contact bookkeeping stays bit-exact).
//
// Reference behaviour followed (paths relative to the reference src/):
//   geometry per triangle      surface_mesh_I.h:302-470, multi_node_mesh_I.h:153-172
//   topology / active flags    surface_mesh_I.h:474-582,1040-1236 (host, setup time, dem_mesh_host.cpp part below)
//   candidate lists            fix_neighlist_mesh.cpp:230-501, tri_mesh_I.h:275-305
//   sphere/triangle contact    tri_mesh_I.h:65-271, math_extra_liggghts.h:583-594
//   contact rows + coplanar    fix_contact_history_mesh_I.h:51-215, fix_contact_history_mesh.cpp:315-505
//   wall force driver          fix_wall_gran.cpp:803-982, fix_wall_gran_base.h:159-367
//   total force on a mesh      mesh_module_stress.cpp:286-345,479-488 (fix mesh/surface/stress)
//   mesh motion                fix_move_mesh.cpp:221-238, mesh_mover_linear.cpp:94-112, multi_node_mesh_I.h:502-526,792-826
//                              rotate: mesh_mover_rotation.cpp:58-125, multi_node_mesh_I.h:620-672, tracking_mesh_I.h:400-411
#pragma once
#include "dem_types.h"

namespace dem {

#define DEM_MAXMESH 8

struct TriRec {  // one triangle, 320 bytes: everything the contact and candidate kernels gather
  double node[9], edgeVec[9], edgeNorm[9], edgeLen[3], surfNorm[3], center[3], rbound;
  int flags;  // [1:0] obtuse node + 1 (0 = none) ; [4:2] edgeActive ; [7:5] cornerActive
  int mesh;   // which `fix mesh/surface` it belongs to
};

struct MeshMeta {
  int atom_type, wall, moving, first, ntri;  // moving: 0 static, 1 `linear`, 2 `rotate`
  double vel[3], precision;
  // rotate: origin, omega*axis, and the per-step quaternion (cos(dphi/2), axis*sin(dphi/2)) evaluated on the host with libm
  double rot_origin[3], rot_omegavec[3], rot_dq[4];
  int rot_trans;  // |origin|^2 > 0 (multi_node_mesh_I.h:648)
  int stress;     // fix mesh/surface/stress: accumulate the total force / torque on this mesh
};

struct MeshP {
  int ntri, nmesh;
  int mslots;  // contact rows per particle
  int mcand;   // candidate rows per particle (ELLPACK width)
  int cap;     // row stride of the per-particle arrays
  int hrec;    // 32-byte history records per contact row (max over the mesh walls)
  TriRec *tri;
  double *nodes_last;   // [9][ntri] node positions at the last rebuild (moving meshes)
  double *mforce;       // [nmesh][6] total force and torque of the current step on the stress-tracking meshes (or null)
  double *mpref;        // [nmesh][3] their reference points (travel with a moving mesh)
  const int *cn;        // coplanar node-neighbours of triangle t: cn[t] .. cn[t+1] index into this same array (CSR: ntri+1 offsets, then the
                        // ascending lists) -- no fixed width: a flat fan of 256 triangles has 255 of them per triangle
  // coarse uniform grid over the box: cell -> ascending triangle ids (CSR)
  const int *cell_start, *cell_tri;
  double gorg[3], ginv[3];
  int gnc[3];
  // per particle: row 0 candidate count, rows 1..mslots partner triangle (-1 = free), then mcand candidate rows
  int *mint;
  double4 *mhist;  // [mslots*hrec][cap]
  MeshMeta meta[DEM_MAXMESH];
  ModelP wm[DEM_MAXMESH];  // wall model of each mesh
  int *overflow;           // [0] candidates needed, [1] contact rows exhausted
};

// launchers implemented in dem_mesh.cu
void mesh_launch_candidates(const MeshP &M, int nlocal, const double4 *xr, double skin, double cdf, cudaStream_t st);
void mesh_launch_step(const StepP &P, const MeshP &M, cudaStream_t st);
void mesh_launch_move(const MeshP &M, int mesh, double dt, double trigsq, int *flag, const int *gate, int gate_mask, cudaStream_t st);
void mesh_launch_hold(const MeshP &M, cudaStream_t st);

}  // namespace dem
