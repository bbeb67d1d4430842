/* dem_b200.h -- C ABI of libdem_b200.so: the B200-native DEM timestep engine.
 *
 * This is the drop-in boundary for the per-timestep particle hot path of
 * LIGGGHTS-INL.  One dem_engine == one GPU == one rank.
 * Every entry point cites the reference interface it replaces (paths relative to
 * the reference tree).  Plain pointers and sizes only; the caller owns every host
 * buffer, the engine owns all device memory, streams and communicators.
 *
 * Error model: every call returns 0 on success or a negative dem_status; the
 * message is available through dem_last_error().  Nothing aborts or throws across
 * the boundary (the reference prints and exit(1)s: src/error.cpp:160-186).
 *
 * There is NO CPU fallback: dem_create fails when no sm_100 device is usable.
 */
#ifndef DEM_B200_H
#define DEM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dem_engine dem_engine;

enum dem_status {
  DEM_OK = 0,
  DEM_ERR_ARG = -1,         /* bad argument / unknown keyword                           */
  DEM_ERR_UNSUPPORTED = -2, /* valid reference syntax that is outside the hot-path scope */
  DEM_ERR_STATE = -3,       /* call order violated (e.g. run before setup)               */
  DEM_ERR_CUDA = -4,        /* CUDA / NCCL runtime failure                               */
  DEM_ERR_OVERFLOW = -5     /* a capacity the engine could not grow                      */
};

/* ---- life cycle ---------------------------------------------------------------------
 * replaces: lammps_open_no_mpi / lammps_close           src/library.h:60-61
 * device   CUDA ordinal; rank/nranks = position in the spatial brick of GPUs;
 * nccl_id  128-byte ncclUniqueId shared by all ranks (NULL when nranks == 1).  With nranks > 1 a NULL id asks,
 *          explicitly, for the communicator that an earlier engine of this process with the same (device, rank,
 *          nranks) released in good standing -- every rank of the job must make the same choice; a non-NULL id
 *          always builds a fresh communicator for exactly the ranks that share the id;
 * stream   cudaStream_t to launch on (NULL = legacy default stream).                  */
int dem_create(dem_engine **out, int device, int rank, int nranks, const void *nccl_id, void *stream);
void dem_destroy(dem_engine *e);
/* multi-GPU plumbing: rank 0 creates the 128-byte ncclUniqueId and shares it (e.g. torch.distributed broadcast)
 * before every rank calls dem_create.  replaces: MPI_Init / MPI_COMM_WORLD of lammps_open  src/library.h:59  */
int dem_nccl_unique_id(void *out128);
/* the brick this rank owns (processor grid, grid position, sub-box): Comm::set_proc_grid  src/comm.cpp,
 * Domain::set_local_box  src/domain.cpp.  Pure host logic, valid after dem_set_box / dem_set_processors.  */
int dem_decomposition(dem_engine *e, int pgrid[3], int myloc[3], double sublo[3], double subhi[3]);
/* The same decomposition without an engine or a GPU (pure host): brick of `rank`, its face neighbours
 * (neigh[2*d], neigh[2*d+1] = lower / upper neighbour in dimension d, -1 = none) and mine[i] = 1 for the positions this
 * rank would keep at dem_upload_particles.  procgrid: {px,py,pz} as `processors`, or NULL / zeros for the engine's choice.
 *                                                   src/comm_brick.cpp:215-330, src/procmap.cpp                  */
int dem_brick_layout(int nranks, int rank, const double lo[3], const double hi[3], const int periodic[3], const int *procgrid,
                     int pgrid[3], int myloc[3], double sublo[3], double subhi[3], int neigh[6],
                     long n, const double *x, int *mine);
const char *dem_last_error(const dem_engine *e);
/* Device blocks released by engines are kept in a process-wide cache for the next engine (set DEM_B200_NO_CACHE=1 to
 * switch it off); this returns them to the driver and reports the bytes freed.                                        */
long dem_trim_memory(void);
const char *dem_version(void);

/* engine knobs that have no deck equivalent (set before dem_setup):
 *   "time_kernels" 0|1   CUDA-event timing of the step kernels, reported by dem_get_stats
 *   "maxneigh" N         initial ELLPACK width of the neighbour rows (default 24; grows on overflow at a rebuild)
 *   "histslots" N        initial history rows per particle (default 16; sized from the contact band at every rebuild)
 *   "cap_factor" x       head room of the per-particle arrays over the uploaded count (1.25; 1.5 with several ranks)
 *   "meshslots" N, "meshcand" N   contact rows / candidate triangles per particle for mesh walls (8 / 16)
 *   "morton" 0|1         Morton (default) or linear cell order of the particle storage
 *   "half_list" 0|1      MEASUREMENT ONLY: half list with fp64 reductions instead of the full list
 *                        without atomics; valid between two rebuilds only, single rank, plain contact models
 *   "debug" bits         profiling aids (skip contact evaluation / list walk / flag all-reduce / halo traffic): timings only */
int dem_set_option(dem_engine *e, const char *name, double value);

/* ---- deck-level settings (same vocabulary as the input script) --------------------- */
/* `units si|cgs|micro`                                   src/update.cpp:160-260        */
int dem_set_units(dem_engine *e, const char *style);
/* `region block` + `create_box` + `boundary p|f|m`       src/domain.cpp, create_box.cpp */
int dem_set_box(dem_engine *e, const double lo[3], const double hi[3], const int periodic[3]);
/* number of atom types of `create_box N`                                                */
int dem_set_ntypes(dem_engine *e, int ntypes);
/* `processors Px Py Pz` (brick of GPUs; product must equal nranks) src/procmap.cpp      */
int dem_set_processors(dem_engine *e, int px, int py, int pz);
/* `neighbor <skin> bin` + `neigh_modify delay D every E check yes|no`
 *                                                        src/neighbor.cpp:1362-1376     */
int dem_set_neighbor(dem_engine *e, double skin, int every, int delay, int check);
/* `timestep dt`                                                                         */
int dem_set_timestep(dem_engine *e, double dt);
/* neigh_modify contact_distance_factor F (neighbor.cpp:1922-1925, F >= 1): list cutoff (r_i + r_j) F + skin, pairs inside the band
 * that do not touch run the models' surfacesClose (pair_gran_base.h:418-422: tangential / rolling history zeroed); the bond
 * models raise it themselves (Neighbor::register_contact_dist_factor keeps the larger value). */
int dem_set_contact_distance_factor(dem_engine *e, double f);
/* `fix ID all property/global <name> scalar|peratomtype|peratomtypepair v...`
 *                                  src/fix_property_global.cpp, global_properties.cpp   */
int dem_set_property(dem_engine *e, const char *name, const char *kind, const double *values, int n);
/* `pair_style gran model hertz|hooke tangential history [cohesion ...]
 *  [rolling_friction cdt|epsd|epsd2] [key on|off ...]` (argv starts at "model")
 *                          src/contact_models.cpp:158-260, src/pair_gran.cpp:229-558    */
int dem_set_pair_style(dem_engine *e, int argc, const char *const *argv);
/* `fix ID all wall/gran model ... primitive type T xplane|yplane|zplane P |
 *  xcylinder|ycylinder|zcylinder R c1 c2 [shear x|y|z v]` (argv starts at "model")
 *                                                        src/fix_wall_gran.cpp:171-330  */
int dem_add_wall_primitive(dem_engine *e, const char *id, int argc, const char *const *argv);
/* `fix ID all mesh/surface[/stress] file F type T [curvature deg] [precision p] [stress on|off] [reference_point x y z]` : the triangles of the file as
 * nodes9[ntri][3][3] (the host layer reads ASCII/binary STL, src/input_mesh_tri.cpp:308-591); load-time
 * scale/move/rotate are applied by the caller.  Geometry, topology and active edge/corner flags follow
 *                          src/surface_mesh_I.h:302-582,1040-1236, src/multi_node_mesh_I.h:153-321  */
int dem_add_mesh(dem_engine *e, const char *id, int atom_type, const double *nodes9, long ntri, int argc, const char *const *argv);
/* `fix ID all move/mesh mesh MESH linear vx vy vz`   src/fix_move_mesh.cpp:221-238, src/mesh_mover_linear.cpp:94-112 */
int dem_move_mesh(dem_engine *e, const char *mesh_id, int argc, const char *const *argv);
/* `fix ID all wall/gran model ... mesh n_meshes N meshes id1 ... [key on|off ...]` (argv starts at "model")
 *      src/fix_wall_gran.cpp:171-330,803-982 ; src/fix_neighlist_mesh.cpp:230-501 ; src/tri_mesh_I.h:65-305 ;
 *      src/fix_contact_history_mesh_I.h:51-215                                                          */
int dem_add_wall_mesh(dem_engine *e, const char *id, int argc, const char *const *argv);
/* `fix ID all gravity g vector x y z`                    src/fix_gravity.cpp:301-383    */
int dem_set_gravity(dem_engine *e, double magnitude, const double dir[3]);
/* `fix ID <group> freeze` : particles whose mask has any bit of groupbit
 *                                                        src/fix_freeze.cpp:121-144     */
int dem_set_freeze(dem_engine *e, int groupbit);
/* `fix ID <group> nve/sphere` : group integrated (default: bit 1 = all)
 *                                                        src/fix_nve_sphere.cpp:134-244 */
int dem_set_integrate(dem_engine *e, int groupbit);

/* ---- particles ------------------------------------------------------------------------
 * replaces: read_data rows `id type diameter density x y z` + velocities
 *           src/atom_vec_sphere.cpp:1055-1083 ; lammps_scatter_atoms src/library.h:72
 * mask may be NULL (all particles in group bit 1), v/omega may be NULL (zero).
 * With nranks > 1 every rank passes the FULL set; each keeps what its brick owns.     */
int dem_upload_particles(dem_engine *e, long n, const int *tag, const int *type, const int *mask,
                         const double *x, const double *v, const double *omega,
                         const double *radius, const double *density);

/* Particles added to a system that already runs: what `create_atoms` and the `fix insert/pack|stream` family do
 * to the path's state (atom.cpp / atom_vec_sphere.cpp create_atom: appended behind the owned atoms, zero force, no contact
 * partners, no wall history).  To be called between two runs; the next dem_setup (== the next `run`'s Verlet::setup) rebuilds the
 * lists and keeps the history of every existing contact.  Several ranks: every rank may pass the whole set, it keeps the
 * particles inside its brick.  Before the first upload it is dem_upload_particles.  Where the positions come from -- the
 * reference's insertion fixes draw them from their own random streams -- is the caller's business. */
int dem_insert_particles(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x,
                         const double *v, const double *omega, const double *radius, const double *density);

/* `fix ID G addforce fx fy fz` (fix_addforce.cpp:234-260, constant components: kind 0, 3 values) and `fix ID G viscous gamma`
 * (fix_viscous.cpp:100-125, one gamma for all types: kind 1, 1 value): per-particle forces added after the pair, gravity and
 * wall forces of a step, in the order of definition (at most 4).  An existing id is replaced; n < 0 removes the fix (`unfix`). */
int dem_set_extra_force(dem_engine *e, const char *id, int kind, int groupbit, const double *values, int n);

/* The timestep in which `fix insert/pack` (fix_insert.cpp:672-905 FixInsert::pre_exchange, fix_insert_pack.cpp:474-597) creates
 * particles, in two halves.  The reference inserts INSIDE a timestep: after the first half step of the existing particles
 * (verlet.cpp:277-286) and before the rebuild it forces (fix->next_reneighbor, neighbor.cpp:1364-1369); the new particles see
 * this step's force evaluation and second half step only.
 *   dem_insert_step_begin: first half step of the particles the engine holds (a no-op while it holds none).  Afterwards
 *                          dem_download "x" returns the positions the reference's overlap check runs against.
 *   dem_insert_step_end:   the n new particles appear (arrays as dem_insert_particles; mass = density * volume with the
 *                          volume rounded as fix_template_sphere.cpp:349-350 forms it), lists are rebuilt, forces evaluated,
 *                          second half step for all.  n == 0: a timestep with a forced rebuild.  Counts as one timestep.
 * The random streams, regions and the overlap search of the insertion are host work of the caller (the deck front end,
 * dem_deck.cpp, restates them). */
int dem_insert_step_begin(dem_engine *e);
int dem_insert_step_end(dem_engine *e, long n, const int *tag, const int *type, const int *mask, const double *x,
                        const double *v, const double *omega, const double *radius, const double *density);

/* ---- run --------------------------------------------------------------------------------
 * dem_setup  == Verlet::setup  (forces with shearupdate = 0)   src/verlet.cpp:134-199
 * dem_run(n) == Verlet::run(n)                                 src/verlet.cpp:264-391   */
int dem_setup(dem_engine *e);
int dem_run(dem_engine *e, long nsteps);

/* ---- read-back (rows ordered by ascending tag over the particles THIS rank owns) ------
 * replaces: lammps_extract_atom / lammps_gather_atoms          src/library.h:66,71
 * field in {"tag","type","mask"} -> int32 x count ; {"radius","rmass","density"} ->
 * double x count ; {"x","v","f","omega","torque"} -> double x 3*count.                  */
long dem_nlocal(const dem_engine *e);
int dem_download(dem_engine *e, const char *field, void *out, long count);
/* granular pair list as the reference's half list: one row per unordered pair whose
 * lower-tag particle is owned by this rank, sorted by (tag_lo,tag_hi); flag != 0 iff the
 * pair holds contact history; hist has dnum doubles per row in the orientation
 * "lower tag is i" (reference: NeighList::firstneigh/firstdouble of listgranhistory,
 * src/neigh_gran.cpp:560-620, src/pair_gran_base.h:213-215).                             */
int dem_pair_count(dem_engine *e, long *npairs, int *dnum);
int dem_download_pairs(dem_engine *e, int *tag_lo, int *tag_hi, int *flag, double *hist);
/* per-particle history of one primitive wall (reference: fix property/atom
 * "history_<wallid>", src/fix_wall_gran.cpp:479-503); rows by ascending tag.            */
int dem_download_wall_history(dem_engine *e, const char *wall_id, double *out, long count);

/* mesh read-back.  field: "nodes" double x 9*ntri (current positions) ; "edge_active","corner_active" int x 3*ntri ;
 * "obtuse","nneighs" int x ntri  (reference: SurfaceMesh::edgeActive/cornerActive/nNeighs, src/surface_mesh.h:152-162) */
int dem_download_mesh(dem_engine *e, const char *mesh_id, const char *field, void *out, long count);
/* per-particle mesh contact rows (reference: FixContactHistoryMesh partner_/contacthistory_,
 * src/fix_contact_history_mesh.h): one row per (particle, triangle) holding history, sorted by (tag, triangle id) */
/* `f_<mesh>[1..9]` of a `fix ID all mesh/surface/stress ...` (dem_add_mesh with "stress on" [, "reference_point x y z"]): total
 * force on the mesh in the last step, total torque about the reference point, the reference point (it travels with a moving
 * mesh).  Rank-local sum with several ranks.   src/mesh_module_stress.cpp:286-345,479-488, src/fix_wall_gran_base.h:350-362 */
int dem_mesh_force(dem_engine *e, const char *mesh_id, double *out9);
int dem_mesh_contact_count(dem_engine *e, const char *mesh_id, long *nrows, int *dnum);
int dem_download_mesh_contacts(dem_engine *e, const char *mesh_id, int *tag, int *tri, double *hist);

typedef struct dem_stats {
  long ntimestep;      /* steps taken since setup                                  */
  long nbuilds;        /* neighbour list builds (reference: neighbor->ncalls)      */
  long nlocal, nghost; /* owned / ghost particles on this rank                     */
  long npairs_full;    /* entries in the full list (both directions)               */
  long ncontacts_full; /* touching entries at the last step (both directions)      */
  long kernel_launches;/* CUDA kernels launched by the engine so far               */
  int maxneigh;        /* ELLPACK width in use                                     */
  int dnum;            /* history doubles per pair                                 */
  double step_kernel_ms;   /* device time of the fused step kernel, last dem_run  */
  long step_kernel_calls;  /* launches that time covers                           */
} dem_stats;
int dem_get_stats(dem_engine *e, dem_stats *out);

/* ---- per-contact output -----------------------------------------------------------
 * replaces: compute ID group pair/gran/local id force torque   src/compute_pair_gran_local.cpp:66-140 (keywords),
 *           :519-640 (post_force_pp: one row per touching pair with the ids and the force / torque on the first particle)
 * Needs dem_set_option("contact_output", 1) before dem_setup.  One row per (owned particle, touching partner) of the last
 * force evaluation that materialised forces (dem_setup, or the last step of dem_run): own tag, partner tag, the force and
 * the torque the pair applies to the owned particle -- a pair of two owned particles gives two rows, one per particle,
 * ordered by (own tag, partner tag).  Plain contact models only.                                                        */
/* compute bond/counter (compute_bond_counter.cpp:101-138, hooks cohesion_model_bond.h:546,1032,1066) as the reference's
 * lammps_extract_compute returns its vector between two runs: out6[0] bonds created since the last call, [1] bonds broken since
 * the last call, [2] "total" = counted + created - broken in unsigned 32-bit arithmetic -- the reference counts existing bonds
 * only on the step after an invocation inside a run (Modify::init resets invoked_vector at every run start), so between runs
 * `counted` is 0 and the entry wraps when more bonds broke than formed; the existing bonds are the set flags of
 * dem_download_pairs.  [3..5] wall bonds: always 0 (out of scope).  Summed over the ranks.  Only the linear bond model feeds
 * the counter (bond/nonlinear looks for a compute style that does not exist, cohesion_model_bond_nonlinear.h:381). */
int dem_bond_counter(dem_engine *e, double *out6);
int dem_contact_count(dem_engine *e, long *n);
int dem_download_contacts(dem_engine *e, int *tag, int *partner, double *force, double *torque);

/* ---- input-script front end (csrc/dem_deck.cpp): the reference's embedding API
 *   void lammps_open_no_mpi(int, char **, void **)   src/library.cpp:85-100   -> dem_create + dem_deck_open
 *   void lammps_file(void *, char *)                 src/library.cpp:130-140  -> dem_deck_file
 *   char *lammps_command(void *, char *)             src/library.cpp:142-160  -> dem_deck_command
 *   void lammps_close(void *)                        src/library.cpp:118-128  -> dem_deck_close + dem_destroy
 * A deck wraps an engine the caller created and still owns.  Commands on the particle hot path (units, atom_style,
 * boundary, newton, communicate, processors, region block, create_box, read_data, neighbor, neigh_modify, group id|type,
 * fix property/global | gravity | wall/gran primitive|mesh | mesh/surface* file [type|scale|move|rotate|curvature|
 * precision] | move/mesh linear|rotate | freeze | nve/sphere, pair_style gran, pair_coeff, timestep, variable NAME
 * equal FORMULA | string|index VALUE (formulas: numbers, + - * / ^, parentheses, v_name, PI, sqrt exp ln log abs sin cos ...),
 * run N [upto]) become the ABI calls above in deck order; output-only commands (thermo*,
 * compute, dump*, ...) are accepted and reported by dem_deck_warnings; everything else returns DEM_ERR_UNSUPPORTED.
 * Errors carry the reference's message texts where one exists (src/input.cpp, src/read_data.cpp). */
typedef struct dem_deck_handle dem_deck;
int dem_deck_open(dem_deck **out, dem_engine *e);
void dem_deck_close(dem_deck *d);
int dem_deck_command(dem_deck *d, const char *line);
int dem_deck_file(dem_deck *d, const char *path);
const char *dem_deck_last_error(const dem_deck *d);
const char *dem_deck_warnings(const dem_deck *d);
long dem_deck_ntimestep(const dem_deck *d);
/* thermo output (thermo.cpp:311-330 header, :334-400 lines: one line at the start of a run, on the multiples of `thermo N` and on
 * the last step; keywords step atoms ke erotate cpu time elapsed, style `one` = step atoms ke cpu) collected so far, and a
 * switch that also prints it to stdout as it is produced (the reference's screen). */
const char *dem_deck_output(const dem_deck *d);
int dem_deck_screen(dem_deck *d, int on);

#ifdef __cplusplus
}
#endif
#endif /* DEM_B200_H */
