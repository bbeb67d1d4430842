#!/bin/sh
# oracle/gen_ref_styles.sh -- TEST INFRASTRUCTURE.
# Writes the style_*.h include lists the reference build expects (what
# `Make.sh style` would produce: one #include per header that mentions the
# registration macro) into $2, reading headers in place from $1 (= reference src).
# Nothing from the reference is copied; the outputs are lists of #include lines.
set -e
SRC="$1"; OUT="$2"; mkdir -p "$OUT"
gen() { # macro prefix name
  : > "$OUT/style_$3.h"
  for f in $(cd "$SRC" && grep -sl "$1" $2*.h | sort); do echo "#include \"$f\"" >> "$OUT/style_$3.h"; done
}
gen ANGLE_CLASS angle_ angle;        gen ATOM_CLASS atom_vec_ atom;      gen BODY_CLASS body_ body
gen BOND_CLASS bond_ bond;           gen COMMAND_CLASS "" command;       gen COMPUTE_CLASS compute_ compute
gen DIHEDRAL_CLASS dihedral_ dihedral; gen DUMP_CLASS dump_ dump;        gen FIX_CLASS fix_ fix
gen IMPROPER_CLASS improper_ improper; gen INTEGRATE_CLASS "" integrate; gen KSPACE_CLASS "" kspace
gen MINIMIZE_CLASS min_ minimize;    gen PAIR_CLASS pair_ pair;          gen SURFACE_MODEL surface_model_ surface_model
gen NORMAL_MODEL normal_model_ normal_model; gen TANGENTIAL_MODEL tangential_model_ tangential_model
gen COHESION_MODEL cohesion_model_ cohesion_model; gen ROLLING_MODEL rolling_model_ rolling_model
gen READER_CLASS reader_ reader;     gen REGION_CLASS region_ region
gen CFD_DATACOUPLING_CLASS cfd_datacoupling_ cfd_datacoupling; gen CFD_REGIONMODEL_CLASS cfd_regionmodel_ cfd_regionmodel
gen LB_CLASS "" lb;                  gen SPH_KERNEL_CLASS sph_kernel_ sph_kernel
gen MESHMODULE_CLASS mesh_module_ mesh_module; gen MESHMOVER_CLASS mesh_mover_ mesh_mover
: > "$OUT/style_contact_model.h"   # no whitelist shipped -> generic (virtual-dispatch) contact model
echo "#define LIGGGHTS_VERSION \"oracle build of $(cat $SRC/version_liggghts_branch.txt) $(cat $SRC/version_liggghts.txt)\"" > "$OUT/version_liggghts.h"
