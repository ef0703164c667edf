#!/bin/bash
# TEST INFRASTRUCTURE -- generator of tests/golden/hctest_lya32 (BASELINE config 1: the reference's own hctest snapshot).
#
# Builds the UNMODIFIED reference (Nyx + vendored AMReX + vendored SUNDIALS, read where they lie under /root/reference) OUT OF TREE
# under $WORK and runs   Exec/LyA inputs.rt max_step=1 nyx.hctest_example_write=1   (Exec/HeatCoolTests/example_setup.sh:37-52,
# SURVEY 8c recipe), which makes Nyx::integrate_state_struct dump hctest/{inputs.0,BADMAP.0,Chunk.0.0}
# (Source/HeatCool/integrate_state_with_source_3d.cpp:82-125, sdc_writeOn f_rhs_struct.H:587-655).  Then the same binary replays
# the snapshot (nyx.hctest_example_read=1, sdc_readFrom f_rhs_struct.H:657-697) with nyx.hctest_example_write=1 on a second
# index, so that the reference's OUTPUT of the replayed step is on disk too: the reference's replay driver compares nothing and writes
# nothing, so the post-step S_new / D_new / IR are captured by a tiny driver compiled against the reference objects
# (tests/golden/hctest_replay_ref.cpp).  Nothing of the reference is copied into the repository; only the generated data files are.
#
#   usage: tests/golden/make_hctest_fixture.sh [WORK=/tmp/nyx_ref_build]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference
WORK="${1:-/tmp/nyx_ref_build}"
SUN=$REF/subprojects/sundials
INST=$WORK/sundials_inst
JOBS=${JOBS:-8}
mkdir -p "$WORK" "$INST/lib" "$INST/include"

# ---- (1) SUNDIALS 6.3.0 static libraries, plain gcc + ar (the CMake superbuild needs .git)
if [ ! -f "$INST/lib/libsundials_cvode.a" ]; then
  cp -r $SUN/include/* "$INST/include/"
  cp "$HERE/../../oracle/shim/sundials/sundials/sundials_config.h" "$HERE/../../oracle/shim/sundials/sundials/sundials_export.h" "$INST/include/sundials/"
  mkdir -p "$WORK/sunobj"
  CF="-O3 -fPIC -fopenmp -I$INST/include -I$SUN/src/sundials -I$SUN/src/cvode"
  gen="sundials_nvector sundials_math sundials_context sundials_nonlinearsolver sundials_linearsolver sundials_matrix sundials_logger sundials_profiler sundials_nvector_senswrapper sundials_futils sundials_memory sundials_version"
  for f in $gen; do /usr/bin/gcc $CF -c $SUN/src/sundials/$f.c -o "$WORK/sunobj/$f.o"; done
  for f in cvode cvode_io cvode_diag cvode_nls cvode_proj cvode_fused_stubs cvode_ls cvode_bandpre cvode_bbdpre cvode_direct cvode_spils; do
    [ -f $SUN/src/cvode/$f.c ] && /usr/bin/gcc $CF -c $SUN/src/cvode/$f.c -o "$WORK/sunobj/$f.o"; done
  /usr/bin/gcc $CF -c $SUN/src/sunnonlinsol/newton/sunnonlinsol_newton.c -o "$WORK/sunobj/sunnonlinsol_newton.o"
  /usr/bin/gcc $CF -c $SUN/src/sunnonlinsol/fixedpoint/sunnonlinsol_fixedpoint.c -o "$WORK/sunobj/sunnonlinsol_fixedpoint.o"
  for d in band dense spgmr spfgmr spbcgs sptfqmr pcg; do /usr/bin/gcc $CF -c $SUN/src/sunlinsol/$d/sunlinsol_$d.c -o "$WORK/sunobj/sunlinsol_$d.o"; done
  for d in band dense sparse; do /usr/bin/gcc $CF -c $SUN/src/sunmatrix/$d/sunmatrix_$d.c -o "$WORK/sunobj/sunmatrix_$d.o"; done
  for f in sundials_band sundials_dense sundials_direct sundials_iterative; do /usr/bin/gcc $CF -c $SUN/src/sundials/$f.c -o "$WORK/sunobj/$f.o"; done
  /usr/bin/gcc $CF -c $SUN/src/nvector/serial/nvector_serial.c -o "$WORK/sunobj/nvector_serial.o"
  /usr/bin/gcc $CF -c $SUN/src/nvector/openmp/nvector_openmp.c -o "$WORK/sunobj/nvector_openmp.o"
  ar rcs "$INST/lib/libsundials_nvecserial.a" "$WORK/sunobj/nvector_serial.o"
  ar rcs "$INST/lib/libsundials_nvecopenmp.a" "$WORK/sunobj/nvector_openmp.o"
  rm "$WORK/sunobj/nvector_serial.o" "$WORK/sunobj/nvector_openmp.o"
  ar rcs "$INST/lib/libsundials_cvode.a" "$WORK"/sunobj/*.o
fi

# ---- (2) the reference executable: Exec/LyA through the reference's GNUmake, out of tree (/root/reference is read-only)
if [ ! -d "$WORK/LyA" ]; then cp -r $REF/Exec/LyA "$WORK/LyA"; fi
cd "$WORK/LyA"
MK="TOP=$REF AMREX_HOME=$REF/subprojects/amrex USE_MPI=FALSE USE_OMP=TRUE USE_ARKODE_LIBS=FALSE SUNDIALS_ROOT=$INST CXX=/usr/bin/g++ CC=/usr/bin/gcc"
if ! ls Nyx3d.*.ex >/dev/null 2>&1; then make -j$JOBS $MK >> "$WORK/make_lya.log" 2>&1 || { tail -30 "$WORK/make_lya.log"; exit 1; }; fi
EXE=$(ls "$WORK"/LyA/Nyx3d.*.ex | head -1)

# ---- (3) the snapshot: first coarse step of inputs.rt (32^3, z ~ 100-160: UVB off)
rm -rf hctest; mkdir -p hctest
OMP_NUM_THREADS=1 "$EXE" inputs.rt max_step=1 nyx.hctest_example_write=1 nyx.v=2 amr.plot_int=-1 amr.check_int=-1 > "$WORK/run_write.log" 2>&1
ls -la hctest
echo "snapshot written by: $EXE"

# ---- (4) the reference's answer for that snapshot: hctest_replay_ref.cpp compiled with the flags of the reference's own build
# (taken from its make log) and linked against its objects (minus the two that hold main)
OBJ="$WORK/LyA/tmp_build_dir/o/3d.gnu.TPROF.OMP.EXE"
# (the link line carries the same -D / -I set as every compile line; it is in the log whether or not objects were rebuilt)
CMD=$(grep -- "-o Nyx3d" "$WORK/make_lya.log" | tail -1)
${CMD%% -Xlinker*} -c "$HERE/hctest_replay_ref.cpp" -o "$WORK/hctest_replay_ref.o"
/usr/bin/g++ -fopenmp -pthread -o "$WORK/hctest_replay_ref.ex" "$WORK/hctest_replay_ref.o" $(ls "$OBJ"/*.o | grep -v -e '/main.o' -e '/nyx_main.o') \
  -L"$INST/lib" -lsundials_cvode -lsundials_nvecserial -lsundials_nvecopenmp
OMP_NUM_THREADS=1 "$WORK/hctest_replay_ref.ex" hctest/inputs.0 nyx.v=2 > "$WORK/run_replay.log" 2>&1 || { tail -20 "$WORK/run_replay.log"; exit 1; }
grep -E "replayed|nst|Nyx::sdc|steps" "$WORK/run_replay.log" | head -8 || true

# ---- (5) the fixture: snapshot + reference output + provenance
OUT="$HERE/hctest_lya32"
rm -rf "$OUT"; mkdir -p "$OUT"
cp hctest/inputs.0 hctest/BADMAP.0 "$OUT/"
( cd hctest && sha256sum inputs.0 BADMAP.0 Chunk.0.0 Chunk.0.out.0 ) > "$OUT/SHA256SUMS"     # of the uncompressed files
xz -9 -c hctest/Chunk.0.0 > "$OUT/Chunk.0.0.xz"; xz -9 -c hctest/Chunk.0.out.0 > "$OUT/Chunk.0.out.0.xz"   # nyx_b200/hctest.py reads .xz transparently
echo "fixture in $OUT"; ls -la "$OUT"
