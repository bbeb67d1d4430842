/* integration/b200_shim.cpp -- see b200_shim.h.  Host code only; everything the GPU does happens behind the C ABI. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "b200_shim.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "modify.h"
#include "neighbor.h"
#include "output.h"
#include "timer.h"
#include "update.h"

#ifdef B200_SHIM_ORACLE
// test build (integration/Makefile, target orc): the same shim bound to the CPU oracle's orc_* entry points, so that
// the binding is covered where there is no GPU.  Test infrastructure only.
extern "C" {
struct dem_engine; struct dem_deck_handle;
#define DEM(n) orc_##n
#define DECK(n) orc_deck_##n
typedef struct dem_deck_handle dem_deck;
int orc_create(struct dem_engine **, int, int, int, const void *, void *); void orc_destroy(struct dem_engine *);
const char *orc_last_error(const struct dem_engine *);
int orc_set_units(struct dem_engine *, const char *); int orc_set_box(struct dem_engine *, const double *, const double *, const int *);
int orc_set_ntypes(struct dem_engine *, int); int orc_set_neighbor(struct dem_engine *, double, int, int, int); int orc_set_timestep(struct dem_engine *, double);
int orc_set_freeze(struct dem_engine *, int); int orc_set_integrate(struct dem_engine *, int);
int orc_upload_particles(struct dem_engine *, long, const int *, const int *, const int *, const double *, const double *, const double *, const double *, const double *);
int orc_setup(struct dem_engine *); int orc_run(struct dem_engine *, long); long orc_nlocal(const struct dem_engine *);
int orc_set_extra_force(struct dem_engine *, const char *, int, int, const double *, int);
int orc_insert_step_begin(struct dem_engine *);
int orc_insert_step_end(struct dem_engine *, long, const int *, const int *, const int *, const double *, const double *, const double *, const double *, const double *);
int orc_download(struct dem_engine *, const char *, void *, long);
int orc_deck_open(dem_deck **, struct dem_engine *); void orc_deck_close(dem_deck *); int orc_deck_command(dem_deck *, const char *);
const char *orc_deck_last_error(const dem_deck *);
}
struct orc_stats_t { long ntimestep, nbuilds, nlocal, nghost, npairs_full, ncontacts_full, kernel_launches; int maxneigh, dnum; double step_kernel_ms; long step_kernel_calls; };
extern "C" int orc_get_stats(struct dem_engine *, orc_stats_t *);
typedef orc_stats_t dem_stats;
#else
extern "C" {
#include "dem_b200.h"
}
#define DEM(n) dem_##n
#define DECK(n) dem_deck_##n
#endif

using namespace LAMMPS_NS;
#define DBG(...) do { if (getenv("B200_SHIM_DEBUG")) { fprintf(stderr, "[b200 shim] " __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

// the deck replay list of a LAMMPS instance (one per instance: the reference keeps other global registries as well)
static std::map<LAMMPS *, std::vector<std::string> > g_replay;

void LAMMPS_NS::b200_remember(LAMMPS *lmp, const char *head, int narg, char **arg)
{
  std::string line(head);
  for (int k = 0; k < narg; k++) { line += " "; line += arg[k]; }
  DBG("remember: %s", line.c_str());
  g_replay[lmp].push_back(line);
}

FixNVESphereB200::FixNVESphereB200(LAMMPS *lmp, int narg, char **arg) : FixNVESphere(lmp, narg, arg)
{
  DBG("nve/sphere/b200: integrate style is %s", update->integrate_style);
  if (strcmp(update->integrate_style, "verlet") == 0) {
    char *a[1] = {(char *)"verlet/b200"};
    update->create_integrate(1, a, NULL);
    DBG("nve/sphere/b200: integrate style is now %s", update->integrate_style);
  }
}

FixAddForceB200::FixAddForceB200(LAMMPS *lmp, int narg, char **arg) : FixAddForce(lmp, narg, arg)
{
  if (narg != 6) error->all(FLERR, "fix addforce/b200: options (region, energy, every) are outside the b200 hot path");
  for (int k = 0; k < 3; k++) {
    if (strstr(arg[3 + k], "v_") == arg[3 + k]) error->all(FLERR, "fix addforce/b200: variable components are outside the b200 hot path");
    b200_values[k] = force->numeric(FLERR, arg[3 + k]);
  }
}
FixViscousB200::FixViscousB200(LAMMPS *lmp, int narg, char **arg) : FixViscous(lmp, narg, arg)
{
  if (narg != 4) error->all(FLERR, "fix viscous/b200: per-type scale factors are outside the b200 hot path");
  b200_values[0] = force->numeric(FLERR, arg[3]);
}

void PairGranB200::settings(int narg, char **arg)
{
  b200_remember(lmp, "pair_style gran", narg, arg);
  // the reference looks its granular pair style up by exact name (force->pair_match("gran", 1), e.g.
  // fix_contact_history.cpp:218): the suffixed style keeps the plain name
  delete[] force->pair_style;
  force->pair_style = new char[5];
  strcpy(force->pair_style, "gran");
  PairGranProxy::settings(narg, arg);
}

VerletB200::VerletB200(LAMMPS *lmp, int narg, char **arg) : Verlet(lmp, narg, arg), eng(NULL), deck(NULL), replayed(0), uploaded_step(-1), holds(false) { DBG("verlet/b200 created"); }

VerletB200::~VerletB200()
{
  if (deck) DECK(close)(deck);
  if (eng) DEM(destroy)(eng);
}

void VerletB200::fail(const char *what)
{
  char msg[768];
  snprintf(msg, sizeof msg, "verlet/b200: %s: %s", what, deck && DECK(last_error)(deck)[0] ? DECK(last_error)(deck) : (eng ? DEM(last_error)(eng) : "no engine"));
  DBG("%s", msg);
  error->all(FLERR, msg);
}

// deck state that lives in objects rather than in remembered commands, then the commands the engine has not seen yet
void VerletB200::sync_settings()
{
  if (comm->nprocs != 1) error->all(FLERR, "verlet/b200: this binding drives one GPU from a serial LIGGGHTS process");
  if (!eng) {
    if (DEM(create)(&eng, 0, 0, 1, NULL, NULL)) fail("dem_create");
    if (DECK(open)(&deck, eng)) fail("dem_deck_open");
    // the box and the number of atom types go through the front end as the deck lines that made them
    char line[512];
    snprintf(line, sizeof line, "units %s", update->unit_style);
    if (DECK(command)(deck, line)) fail(line);
    snprintf(line, sizeof line, "boundary %s %s %s", domain->xperiodic ? "p" : "f", domain->yperiodic ? "p" : "f", domain->zperiodic ? "p" : "f");
    if (DECK(command)(deck, line)) fail(line);
    snprintf(line, sizeof line, "region b200box block %.17g %.17g %.17g %.17g %.17g %.17g units box", domain->boxlo[0], domain->boxhi[0],
             domain->boxlo[1], domain->boxhi[1], domain->boxlo[2], domain->boxhi[2]);
    if (DECK(command)(deck, line)) fail(line);
    snprintf(line, sizeof line, "create_box %d b200box", atom->ntypes);
    if (DECK(command)(deck, line)) fail(line);
  }
  if (DEM(set_neighbor)(eng, neighbor->skin, neighbor->every, neighbor->delay, neighbor->dist_check)) fail("neighbor");
  if (DEM(set_timestep)(eng, update->dt)) fail("timestep");
  std::vector<std::string> &cmds = g_replay[lmp];
  for (; replayed < cmds.size(); replayed++) {
    DBG("replay: %s", cmds[replayed].c_str());
    if (DECK(command)(deck, cmds[replayed].c_str())) fail(cmds[replayed].c_str());
  }
  // group bits of the fixes that carry no other parameter; every other fix must be one the engine knows, an internal helper
  // of those, or output only -- anything else would silently change the physics
  // (insert/pack, particletemplate/sphere, particledistribution/*: the reference's own insertion fixes keep drawing the particles;
  // VerletB200::insertion_step hands what they create to the engine inside the timestep.  insert/stream is NOT in the list: it
  // moves its particles itself until they have left the insertion face, which the engine does not know)
  static const char *known[] = {"wall/gran", "mesh/surface", "move/mesh", "gravity", "property/global", "property/atom", "contacthistory",
                                "neighlist/mesh", "check/timestep/gran", "print", "ave/", "store", "contactproperty", "insert/pack", "particletemplate/sphere", "STORE",

                                "particledistribution/", NULL};
  // fix addforce / viscous: drop the ones the deck has unfixed, (re)send the others in the order of their definition
  for (size_t k = 0; k < sent_xf.size(); k++) if (modify->find_fix(sent_xf[k].c_str()) < 0) DEM(set_extra_force)(eng, sent_xf[k].c_str(), 0, 0, NULL, -1);
  sent_xf.clear();
  for (int i = 0; i < modify->nfix; i++) {
    Fix *f = modify->fix[i];
    if (FixAddForceB200 *a = dynamic_cast<FixAddForceB200 *>(f)) { if (DEM(set_extra_force)(eng, f->id, 0, f->groupbit, a->b200_values, 3)) fail("addforce"); sent_xf.push_back(f->id); continue; }
    if (FixViscousB200 *v = dynamic_cast<FixViscousB200 *>(f)) { if (DEM(set_extra_force)(eng, f->id, 1, f->groupbit, v->b200_values, 1)) fail("viscous"); sent_xf.push_back(f->id); continue; }
    if (strcmp(f->style, "freeze") == 0) { if (DEM(set_freeze)(eng, f->groupbit)) fail("freeze"); continue; }
    if (strcmp(f->style, "nve/sphere") == 0) { if (DEM(set_integrate)(eng, f->groupbit)) fail("nve/sphere"); continue; }
    bool ok = false;
    for (int k = 0; known[k]; k++) if (strncmp(f->style, known[k], strlen(known[k])) == 0) ok = true;
    if (!ok) { char msg[256]; snprintf(msg, sizeof msg, "verlet/b200: fix style '%s' (fix %s) is outside the b200 hot path", f->style, f->id); error->all(FLERR, msg); }
  }
}

void VerletB200::push_state()
{
  const int n = atom->nlocal;
  if (DEM(upload_particles)(eng, n, atom->tag, atom->type, atom->mask, n ? &atom->x[0][0] : NULL, n ? &atom->v[0][0] : NULL,
                             n ? &atom->omega[0][0] : NULL, atom->radius, atom->density)) fail("dem_upload_particles");
  uploaded_step = update->ntimestep;
  holds = true;
}

// engine -> atom arrays (the engine returns fields ordered by tag; the reference's local order is its own)
void VerletB200::pull_state(bool forces)
{
  const long n = DEM(nlocal)(eng);
  if (n == 0 && atom->nlocal == 0) return;
  if (n != atom->nlocal) error->all(FLERR, "verlet/b200: particle count changed");
  std::vector<int> tags(n), order(n);
  if (DEM(download)(eng, "tag", tags.data(), n)) fail("download tag");
  std::map<int, int> where;
  for (int i = 0; i < atom->nlocal; i++) where[atom->tag[i]] = i;
  for (long k = 0; k < n; k++) order[k] = where[tags[k]];
  std::vector<double> buf(3 * (size_t)n);
  struct { const char *name; double **dst; } fields[] = {{"x", atom->x}, {"v", atom->v}, {"omega", atom->omega}, {"f", atom->f}, {"torque", atom->torque}};
  for (auto &fd : fields) {
    if (!forces && (fd.dst == atom->f || fd.dst == atom->torque)) continue;
    if (DEM(download)(eng, fd.name, buf.data(), n)) fail(fd.name);
    for (long k = 0; k < n; k++) { double *d = fd.dst[order[k]]; d[0] = buf[3 * k]; d[1] = buf[3 * k + 1]; d[2] = buf[3 * k + 2]; }
  }
}

// Verlet::setup (verlet.cpp:134-199) runs first, unchanged: the reference builds its own lists and forces once, which keeps
// its output machinery (thermo, dumps, computes) valid; then the engine is brought to the same point
void VerletB200::setup()
{
  DBG("setup: base");
  Verlet::setup();
  DBG("setup: sync_settings");
  sync_settings();
  DBG("setup: push/setup/pull");
  if (atom->nlocal == 0 && !holds) return;                // an empty box that an insertion fix will fill: nothing to set up yet
  if (uploaded_step != update->ntimestep) push_state();   // first run, or the deck changed the particles between two runs
  if (DEM(setup)(eng)) fail("dem_setup");
  pull_state();
}

// the next timestep on which one of the reference's insertion fixes acts (Fix::next_reneighbor, fix_insert.cpp:393-395,897-899)
static bigint next_insertion(Modify *modify, bigint now)
{
  bigint next = -1;
  for (int i = 0; i < modify->nfix; i++) {
    Fix *f = modify->fix[i];
    if (strcmp(f->style, "insert/pack") == 0 && f->force_reneighbor && f->next_reneighbor > now && (next < 0 || f->next_reneighbor < next)) next = f->next_reneighbor;
  }
  return next;
}

// One timestep in which the reference's OWN insertion fixes create particles (FixInsert::pre_exchange, fix_insert.cpp:672-905,
// between the first half step and the forced rebuild): the engine does the first half step, the atom arrays receive the
// positions the fixes check overlaps against, the fixes run unchanged -- their random streams, regions and templates are the
// reference's -- and whatever they appended to the atom arrays enters the engine for the second half of the step.
void VerletB200::insertion_step()
{
  update->ntimestep++;
  if (DEM(insert_step_begin)(eng)) fail("dem_insert_step_begin");
  if (holds) pull_state(false);
  const int n0 = atom->nlocal;
  atom->nghost = 0;  // (ghosts of the last build are stale; one process, no periodic images in the overlap check)
  for (int i = 0; i < modify->nfix; i++)
    if (strcmp(modify->fix[i]->style, "insert/pack") == 0) modify->fix[i]->pre_exchange();
  const int nnew = atom->nlocal - n0;
  DBG("insertion step %ld: %d new particles", (long)update->ntimestep, nnew);
  if (DEM(insert_step_end)(eng, nnew, atom->tag + n0, atom->type + n0, atom->mask + n0, nnew ? &atom->x[n0][0] : NULL, nnew ? &atom->v[n0][0] : NULL,
                           nnew ? &atom->omega[n0][0] : NULL, atom->radius + n0, atom->density + n0)) fail("dem_insert_step_end");
  if (nnew) holds = true;
  uploaded_step = update->ntimestep;
  if (holds) pull_state();
}

void VerletB200::run(int n)
{
  bigint left = n;
  while (left > 0) {
    bigint m = left;
    if (output->next > update->ntimestep && output->next - update->ntimestep < m) m = output->next - update->ntimestep;
    const bigint ins = next_insertion(modify, update->ntimestep);
    if (ins == update->ntimestep + 1) { insertion_step(); m = 1; }
    else {
      if (ins > 0 && ins - 1 - update->ntimestep < m) m = ins - 1 - update->ntimestep;
      if (holds) { if (DEM(run)(eng, (long)m)) fail("dem_run"); }  // (an empty box has nothing to step)
      update->ntimestep += m;
      if (holds) pull_state();
      uploaded_step = update->ntimestep;
    }
    left -= m;
    if (update->ntimestep == output->next) {
      // thermo keywords such as pe check that energies were tallied on this step (compute_pe.cpp:105-106); the granular
      // styles tally none, Verlet::run would have marked the step through ev_set (integrate.cpp:117-150)
      update->eflag_global = update->vflag_global = update->ntimestep;
      timer->stamp();
      output->write(update->ntimestep);
      timer->stamp(TIME_OUTPUT);
    }
  }
  dem_stats st;
  if (DEM(get_stats)(eng, &st) == 0) neighbor->ncalls = (int)st.nbuilds;
}
