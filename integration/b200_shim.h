/* integration/b200_shim.h -- the `-suffix b200` binding of libdem_b200.so inside the reference tree.
 *
 * What a LIGGGHTS-INL maintainer adds to src/ (here it is compiled into a copy of the reference build by
 * integration/Makefile, the reference sources stay where they are): with `-suffix b200` on the command line every
 * factory of the reference tries `<style>/b200` first (src/force.cpp:227-236, src/modify.cpp:837-845,
 * src/update.cpp:374-401).  The shim styles below are the reference's own classes -- each one only remembers the
 * words of its deck command and behaves like its base class otherwise, so thermo output, dumps and restart files keep
 * working -- plus one integrate style, `verlet/b200`, that replays the remembered commands into a dem_engine through
 * the input-script front end of the C ABI (include/dem_b200.h: dem_deck_command) and hands `run N` to dem_run().
 * An unmodified input deck therefore runs on the GPU engine with `lmp -suffix b200 -in deck`.
 *
 * The style registration macros are collected by `Make.sh style` (src/Make.sh:80-105) like those of any other header. */
#ifdef FIX_CLASS

FixStyle(wall/gran/b200,FixWallGranB200)
FixStyle(mesh/surface/b200,FixMeshSurfaceB200)
FixStyle(mesh/surface/stress/b200,FixMeshSurfaceB200)
FixStyle(move/mesh/b200,FixMoveMeshB200)
FixStyle(gravity/b200,FixGravityB200)
FixStyle(property/global/b200,FixPropertyGlobalB200)
FixStyle(nve/sphere/b200,FixNVESphereB200)
FixStyle(addforce/b200,FixAddForceB200)
FixStyle(viscous/b200,FixViscousB200)

#elif defined(PAIR_CLASS)

PairStyle(gran/b200,PairGranB200)

#elif defined(INTEGRATE_CLASS)

IntegrateStyle(verlet/b200,VerletB200)

#else

#ifndef LMP_B200_SHIM_H
#define LMP_B200_SHIM_H

#include <string>
#include <vector>
#include "verlet.h"
#include "pair_gran_proxy.h"
#include "fix_wall_gran.h"
#include "fix_mesh_surface.h"
#include "fix_move_mesh.h"
#include "fix_gravity.h"
#include "fix_property_global.h"
#include "fix_nve_sphere.h"
#include "fix_addforce.h"
#include "fix_viscous.h"

struct dem_engine;        // include/dem_b200.h (opaque)
struct dem_deck_handle;

namespace LAMMPS_NS {

void b200_remember(class LAMMPS *, const char *head, int narg, char **arg);  // appends "<head> arg0 arg1 ..." to the deck replay list

#define B200_FIX_SHIM(Shim, Base)                                                                                  \
  class Shim : public Base {                                                                                       \
   public:                                                                                                         \
    Shim(class LAMMPS *lmp, int narg, char **arg) : Base(lmp, narg, arg) { b200_remember(lmp, "fix", narg, arg); } \
  };
B200_FIX_SHIM(FixWallGranB200, FixWallGran)
B200_FIX_SHIM(FixMeshSurfaceB200, FixMeshSurface)
B200_FIX_SHIM(FixMoveMeshB200, FixMoveMesh)
B200_FIX_SHIM(FixGravityB200, FixGravity)
B200_FIX_SHIM(FixPropertyGlobalB200, FixPropertyGlobal)
// fix addforce / fix viscous act on a group of the reference: their b200 variants keep the numbers of their command, which
// VerletB200::sync_settings hands to dem_set_extra_force together with the group's bit
class FixAddForceB200 : public FixAddForce {
 public:
  FixAddForceB200(class LAMMPS *lmp, int narg, char **arg);
  double b200_values[3];
};
class FixViscousB200 : public FixViscous {
 public:
  FixViscousB200(class LAMMPS *lmp, int narg, char **arg);
  double b200_values[1];
};

// `fix nve/sphere` is part of every granular deck: its b200 variant also switches the integrator, because the default
// `verlet` is created before any deck command is read and never sees the suffix (src/update.cpp:103)
class FixNVESphereB200 : public FixNVESphere {
 public:
  FixNVESphereB200(class LAMMPS *, int, char **);
};

class PairGranB200 : public PairGranProxy {
 public:
  PairGranB200(class LAMMPS *lmp) : PairGranProxy(lmp) {}
  virtual void settings(int narg, char **arg);
};

class VerletB200 : public Verlet {
 public:
  VerletB200(class LAMMPS *, int, char **);
  virtual ~VerletB200();
  virtual void setup();
  virtual void run(int);

 private:
  ::dem_engine *eng;
  ::dem_deck_handle *deck;
  size_t replayed;       // commands of the replay list the engine has seen
  bigint uploaded_step;  // timestep at which the engine received the particle state
  void fail(const char *what);
  void sync_settings();
  std::vector<std::string> sent_xf;  // ids of the fix addforce / viscous the engine knows
  bool holds;            // the engine holds particles (false while an empty box waits for an insertion fix)
  void push_state();
  void pull_state(bool forces = true);
  void insertion_step();
};

}  // namespace LAMMPS_NS

#endif
#endif
