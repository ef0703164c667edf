#!/bin/bash
# A16: the drop-in compiled against the reference's REAL headers and linked with the reference's UNMODIFIED objects.
#
# Needs the out-of-tree reference build of tests/golden/make_hctest_fixture.sh ($WORK, default /tmp/nyx_ref_build).  Compiles
# nyx_b200/csrc/nyx_heatcool_dropin.cpp with the flag set of the reference's own GNUmake build (CPU AMReX, OpenMP: the _host entry points)
# and links two executables into tests/_build/real/ (git-ignored; they travel to the GPU box):
#   Nyx3d.dropin.ex          the Exec/LyA executable with integrate_state_vec_3d.o and integrate_state_with_source_3d.o REPLACED by the
#                            drop-in, every other object (strang_reactions.o, sdc_reactions.o, Nyx_advance.o, Nyx_setup.o, ...) as the
#                            reference's build made it -- what INTEGRATION.md section 1 prescribes for Make.package;
#   hctest_replay.dropin.ex  tests/golden/hctest_replay_ref.cpp (the Exec/HeatCoolTests replay + a dump of the result) over the same objects.
set -eu
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
WORK="${1:-/tmp/nyx_ref_build}"
OBJ="$WORK/LyA/tmp_build_dir/o/3d.gnu.TPROF.OMP.EXE"
[ -d "$OBJ" ] || { echo "no reference build under $WORK (run tests/golden/make_hctest_fixture.sh)"; exit 2; }
OUT="$HERE/_build/real"; mkdir -p "$OUT"
CMD=$(grep -- "-o Nyx3d" "$WORK/make_lya.log" | tail -1)
CXXFLAGS="${CMD%% -Xlinker*}"
( cd "$WORK/LyA" && $CXXFLAGS -I"$ROOT/include" -c "$ROOT/nyx_b200/csrc/nyx_heatcool_dropin.cpp" -o "$WORK/nyx_heatcool_dropin.o" )
# Nyx::update_state_with_sources (floor and conservative enforce_minimum_density) in place of the reference's translation unit
( cd "$WORK/LyA" && $CXXFLAGS -I"$ROOT/include" -c "$ROOT/nyx_b200/csrc/nyx_sources_dropin.cpp" -o "$WORK/nyx_sources_dropin.o" )
[ -f "$WORK/hctest_replay_ref.o" ] || ( cd "$WORK/LyA" && $CXXFLAGS -c "$HERE/golden/hctest_replay_ref.cpp" -o "$WORK/hctest_replay_ref.o" )
# the non-SDC signature of Nyx::update_state_with_sources (no reset_e_src argument): syntax check against the real headers without -DSDC
( cd "$WORK/LyA" && ${CXXFLAGS// -DSDC/} -I"$ROOT/include" -fsyntax-only "$ROOT/nyx_b200/csrc/nyx_sources_dropin.cpp" )
KEEP=$(ls "$OBJ"/*.o | grep -v -e '/integrate_state_vec_3d.o' -e '/integrate_state_with_source_3d.o')
LIBS="-L$WORK/sundials_inst/lib -lsundials_cvode -lsundials_nvecserial -lsundials_nvecopenmp -L$ROOT/nyx_b200/csrc -lnyx_hc -Wl,-rpath,/root/repo/nyx_b200/csrc -Wl,-rpath,$ROOT/nyx_b200/csrc"
/usr/bin/g++ -fopenmp -pthread -o "$OUT/Nyx3d.dropin.ex" $KEEP "$WORK/nyx_heatcool_dropin.o" $LIBS
#   Nyx3d.dropin_src.ex      the same, and Nyx_update_state_with_sources.o REPLACED by nyx_sources_dropin.o as well
/usr/bin/g++ -fopenmp -pthread -o "$OUT/Nyx3d.dropin_src.ex" $(echo "$KEEP" | grep -v '/Nyx_update_state_with_sources.o') "$WORK/nyx_heatcool_dropin.o" "$WORK/nyx_sources_dropin.o" $LIBS
strip "$OUT/Nyx3d.dropin_src.ex"
/usr/bin/g++ -fopenmp -pthread -o "$OUT/hctest_replay.dropin.ex" "$WORK/hctest_replay_ref.o" $(echo "$KEEP" | grep -v -e '/main.o' -e '/nyx_main.o') "$WORK/nyx_heatcool_dropin.o" $LIBS
strip "$OUT/Nyx3d.dropin.ex" "$OUT/hctest_replay.dropin.ex"
# SAVE_REACT flavour (needs the second reference build of tests/golden/make_react_fixture.sh): the drop-in compiled with -DSAVE_REACT against
# the objects of the reference built with USE_SAVE_REACT=TRUE
if [ -d "$WORK/LyA_react/tmp_build_dir/o/3d.gnu.TPROF.OMP.EXE" ]; then
  OBJR="$WORK/LyA_react/tmp_build_dir/o/3d.gnu.TPROF.OMP.EXE"
  CMDR=$(grep -- "-o Nyx3d" "$WORK/make_lya_react.log" | tail -1)
  ( cd "$WORK/LyA_react" && ${CMDR%% -Xlinker*} -I"$ROOT/include" -c "$ROOT/nyx_b200/csrc/nyx_heatcool_dropin.cpp" -o "$WORK/nyx_heatcool_dropin_react.o" )
  KEEPR=$(ls "$OBJR"/*.o | grep -v -e '/integrate_state_vec_3d.o' -e '/integrate_state_with_source_3d.o')
  /usr/bin/g++ -fopenmp -pthread -o "$OUT/Nyx3d.dropin_react.ex" $KEEPR "$WORK/nyx_heatcool_dropin_react.o" $LIBS
  strip "$OUT/Nyx3d.dropin_react.ex"
fi
# the reference's own executable beside them (oracle/_ref: built from the reference sources, test infrastructure)
mkdir -p "$ROOT/oracle/_ref"
cp "$(ls "$WORK"/LyA/Nyx3d.*.ex | head -1)" "$ROOT/oracle/_ref/Nyx3d.reference.ex"; strip "$ROOT/oracle/_ref/Nyx3d.reference.ex"
cp "$WORK/hctest_replay_ref.ex" "$ROOT/oracle/_ref/hctest_replay.reference.ex"; strip "$ROOT/oracle/_ref/hctest_replay.reference.ex"
# what the LyA run needs beside the executable (inputs.rt, the 32^3 initial conditions, the UVB table): staged next to it
cp "$WORK/LyA/inputs.rt" "$WORK/LyA/32.nyx" "$WORK/LyA/TREECOOL_middle" "$OUT/"
nm -D --undefined-only "$OUT/Nyx3d.dropin.ex" | grep -c " hc_" || true | sed 's/^/C-ABI symbols imported by Nyx3d.dropin.ex: /'
ls -la "$OUT"
