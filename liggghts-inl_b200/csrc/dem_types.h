// dem_types.h -- plain structs shared by host orchestration and device kernels.
#pragma once
#include <cuda_runtime.h>

namespace dem {

enum { N_OFF = 0, N_HERTZ = 1, N_HOOKE = 2, N_HYST1 = 3, N_HYST2 = 4 };  // N_HYST*: INL laws hysteretic/nonlinear1|2 (pair style only)
enum { R_OFF = 0, R_CDT = 1, R_EPSD = 2, R_EPSD2 = 3 };
// per-type-pair tables, each (ntypes+1)^2 doubles, concatenated in this order
enum { T_YEFF = 0, T_GEFF, T_BETA, T_CORLOG, T_MU, T_RMU, T_RVISC, T_SQ2Y, T_SQ8G, T_INV8G,
       // bond models (cohesion_model_bond.h:76-178, cohesion_model_bond_nonlinear.h:77-147)
       T_B_LAMBDA, T_B_KN, T_B_KT, T_B_DFN, T_B_DFT, T_B_DTN, T_B_DTT, T_B_MAXDIST, T_B_MAXSIGMA, T_B_MAXTAU, T_B_CREATEDIST, T_B_RATIOTC,
       T_B_K_FN1, T_B_KU_FN1, T_B_KC_FN1, T_B_K_FN2, T_B_KU_FN2, T_B_KC_FN2, T_B_K_FT, T_B_K_TN, T_B_KU_TN, T_B_KC_TN, T_B_K_TT, T_B_KU_TT, T_B_KC_TT,
       // normal models hysteretic/nonlinear1|2 (normal_model_hysteretic_nonlinear1.h:105-133)
       T_H_KEL, T_H_KN2K1, T_H_KN2KC, T_H_PHIF, T_H_FADH, T_H_ALPHA, T_H_CIN, T_H_A1, T_H_A2, T_H_A3, T_H_KCIN,
       // linear bond, option dissipationBond: 1 / dissipation{Normal,Tangential}{Force,Torque}Bond (inverted like createDissipationMatrix, global_properties.cpp:1138-1163)
       T_D_FN, T_D_FT, T_D_TN, T_D_TT,
       T_COUNT };
enum { C_OFF = 0, C_BOND = 1, C_BONDNL = 2 };

#define DEM_MAXW 16  // primitive walls per engine (candidate + valid bits share one 32-bit word)

// neighbour word: [31]    partner is the "first" body of the pair (partner tag < own tag)
//                 [30:25] history slot + 1 of this pair in the particle's compact history rows;
//                         0 = pair holds no history (== reference contact_flag == 0)
//                 [24:0]  partner index (owned or ghost), cf. NEIGHMASK lmptype.h:86
// numneigh word:  [15:0] list length, [31:16] history slots in use by this particle
#define NBR_JFIRST 0x80000000u
#define NBR_HIST 0x7E000000u
#define NBR_SLOT_SHIFT 25
#define NBR_IDX 0x01FFFFFFu
#define NBR_MAXSLOTS 62

struct ModelP {
  int normal, tangential, rolling;
  int tdamp, limitForce, torsion, ktToKn;
  int cdtnl2;  // rolling_friction cdtnonlinear2 (rolling_model_cdtnonlinear2.h): the CDT law with the full normal force Fn instead of kn*deltan
  int dnum, off_shear, off_roll;  // reference layout of a history row (fix_contact_history)
  int hrec, rec_shear, rec_roll;   // device layout: hrec 32-byte records per contact, one per sub-model
  int off_norm, rec_norm;          // hysteretic/nonlinear1|2: 12 history doubles of the normal model (3 records), first in the row
  // cohesion bond | bond/nonlinear: nbond history doubles (14 | 28) in records rec_bond.. , followed by one "sticky flag"
  // double (the reference's contact_flags != 0 after a touch or a kept rebuild, see dem_kernels.cuh k_step_bond)
  int cohesion, off_bond, nbond, rec_bond, nbrec;
  int stressBreak, tension, compression, shearf, ntorque, ttorque, createAlways, damping, dampingSmooth, ratioTC;
  int dissipation;  // cohesion bond: dissipationBond on (cohesion_model_bond.h:270,677-687,731,763,798)
};

struct WallP {
  ModelP m;
  int wtype;  // 0..2 plane x,y,z ; 3..5 cylinder x,y,z
  int atom_type;
  int shear, shearDim, shearAxis;
  int hist_row;  // first row of this wall in the whist array
  double param[3];
  double vshear, axisVec[3];
};

enum { MODE_SETUP = 0, MODE_STEP = 1, MODE_LAST = 2 };

#define DEM_MAXRANKS 16
#define FBOX_SERIAL (2 * DEM_MAXRANKS * 4)   // int offset of the serial words in a flag box: [2 slots][DEM_MAXRANKS][4] flags, then [DEM_MAXRANKS] serials
#define FBOX_INTS (FBOX_SERIAL + DEM_MAXRANKS)
// Fused ghost push (several ranks, one decomposed dimension; dem_engine.cu fused_halo_setup): between two rebuilds the step
// kernel itself stores every COPY of a particle's new records -- the ghost slots on the neighbour ranks over NVLink peer
// memory and the periodic images on its own rank -- from its epilogue.  The parameters below are constant between two rebuilds
// and live in device memory.
struct ImgP {
  double4 *bx[3][2], *bv[3][2], *bw[3][2];  // [target][buffer]: 0 this rank, 1 the rank swap 0 sends to, 2 the rank swap 1 sends to
  int pcur0[3];                              // the target's buffer parity at the rebuild (target 0: cur0)
  double prd[3];
};

// per-particle forces of two more post_force fixes, applied in the order of their definition after walls and gravity:
// kind 0 fix addforce (fix_addforce.cpp:234-260, constant components), kind 1 fix viscous (fix_viscous.cpp:100-125, one gamma)
#define DEM_MAXXF 4
struct XForce { int kind, bit; double v[3]; };

struct StepP {
  int nlocal, nall, cap, maxk;
  int lcap;  // row stride of the ELLPACK arrays (nbr, hist)
  // particle records, 32 B each: (x,y,z,radius) (vx,vy,vz,mass) (wx,wy,wz,bits(type|mask<<8))
  const double4 *xr, *vm, *wt;
  double4 *xr_o, *vm_o, *wt_o;
  double4 *xh;  // (xhold, bits(wall candidate | wall-history-valid << 16))
  unsigned *nbr;
  int *numneigh;
  double4 *hist;  // [hslots*hrec][lcap] 32-byte records: contact slot c of particle i at rows c*hrec..
  int hslots;
  double *whist;  // [sum wall dnum][cap]
  double *f, *tq; // [3][cap]
  // owner list (option owner_list, dem_pairs.cuh): per-contact result records [hslots][lcap][2], stamped with the launch serial
  double4 *res;
  double serial;
  double4 *cout;   // option contact_output: force / torque of every evaluated contact [hslots][lcap][2], stamped like res; else null
  const WallP *walls;
  int nwalls;
  int nwc, nwcap;     // primitive-wall candidates (compact list) and row stride of fw
  const int *wlist;   // [nwc] particle index
  double *fw;         // [6][nwcap] wall force / torque of candidate c
  ModelP pm;
  const double *tab;
  int nt1;  // ntypes+1
  double t1[T_B_LAMBDA];  // the contact tables' single entry when ntypes == 1
  long ntimestep, tsCreateBond;  // bond creation step (update->ntimestep == tsCreateBond)
  double dt, dtv, dtf, dtfrot, nktv2p, charVel, cdf, cdfsq, trigsq, cutneighmax;
  double g[3];
  int have_g, have_pair, freezebit, integbit;
  int mode;
  int debug;  // profiling aid (option "debug"): bit0 skip contact evaluation, bit1 skip list walk
  int *flag;  // device flags written by this step: [0] rebuild trigger, [1] history-slot overflow, [2] moving-mesh trigger
  // speculative launch: the flags the PREVIOUS step produced (all-reduced over the ranks).  If a masked entry is set the
  // neighbour list must be rebuilt before this step, so every kernel of the step returns at once and the host relaunches it
  // after the rebuild; the host therefore never has to wait for a step before queueing the next one.
  const int *gate;
  int gate_mask;  // bit0: gate[0] (distance check due this step), bit2: gate[2] (a moving mesh forces the rebuild)
  unsigned long long *ncontact;  // optional counter of touching entries (stats), may be null
  unsigned long long *bondc;     // compute bond/counter: [0] bonds created, [1] bonds broken (each pair counted by its lower-tag particle)
  // fused ghost push (null / unused otherwise): image table of the owned particles (img_first[i]: first entry or -1; entry =
  // (slot | target << 28, shift code, next entry or -1, 0)) and this launch's buffer parity
  const ImgP *img; const int *img_first; const int4 *img_tab;
  int img_par;
  int nxf; XForce xf[DEM_MAXXF];
};

}  // namespace dem
