#!/bin/bash
# TEST INFRASTRUCTURE -- generator of tests/golden/hctest_lya32/react_reference.json (the SAVE_REACT dumps of the reference for the step of
# BASELINE config 1).
#
# Builds the UNMODIFIED reference a second time with USE_SAVE_REACT=TRUE (Exec/Make.Nyx:51-52; out of tree, next to the build of
# make_hctest_fixture.sh, whose SUNDIALS libraries it reuses) and runs the same   Exec/LyA inputs.rt max_step=1 nyx.hctest_example_write=1
# step: Nyx::integrate_state_struct then writes plt_react_in / plt_react_out / plt_react_out_work (integrate_state_with_source_3d.cpp:165-182)
# next to the hctest snapshot.  (The standalone replay driver cannot do this: the SAVE_REACT code needs the level's grids, step counter
# and state times, i.e. a fully initialised Amr object.)  The snapshot of this run has the same valid-region data as the committed one
# (checked below); the three 32^3 FABs are reduced to one SHA-256 per component plus the run's exact a, a_end, dt.
#   usage: tests/golden/make_react_fixture.sh [WORK=/tmp/nyx_ref_build]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference
WORK="${1:-/tmp/nyx_ref_build}"
INST=$WORK/sundials_inst
JOBS=${JOBS:-8}
[ -f "$INST/lib/libsundials_cvode.a" ] || { echo "run make_hctest_fixture.sh first (SUNDIALS libraries)"; exit 1; }
mkdir -p "$WORK/LyA_react"; cd "$WORK/LyA_react"
[ -f GNUmakefile ] || cp -r $REF/Exec/LyA/* .
MK="TOP=$REF AMREX_HOME=$REF/subprojects/amrex USE_MPI=FALSE USE_OMP=TRUE USE_ARKODE_LIBS=FALSE USE_SAVE_REACT=TRUE SUNDIALS_ROOT=$INST CXX=/usr/bin/g++ CC=/usr/bin/gcc"
if ! ls Nyx3d.*.ex >/dev/null 2>&1; then make -j$JOBS $MK >> "$WORK/make_lya_react.log" 2>&1 || { tail -30 "$WORK/make_lya_react.log"; exit 1; }; fi
EXE=$(ls "$WORK"/LyA_react/Nyx3d.*.ex | head -1)
rm -rf hctest plt_react_*; mkdir -p hctest
OMP_NUM_THREADS=1 "$EXE" inputs.rt max_step=1 nyx.hctest_example_write=1 nyx.v=2 amr.plot_int=-1 amr.check_int=-1 > "$WORK/run_write_react.log" 2>&1
python "$HERE/react_reference_digest.py" "$WORK/LyA_react" "$HERE/hctest_lya32"
