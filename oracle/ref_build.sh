#!/bin/bash
# Builds oracle/_ref/libpicnic_ref.so from the REFERENCE's own Chombo-free arithmetic sources, where
# they lie under /root/reference, against the mock headers in oracle/chombo_mock/.  Test
# infrastructure only: the result pins oracle/ (tests/test_ref_pin.py) and is never linked by the
# product.  Nothing is copied out of /root/reference; outputs go to oracle/_ref/ only.
# The rest of the path (MeshInterp*.ChF gathers/deposits, PicChargedSpecies.cpp, scattering models)
# needs Chombo proper + chfpp + gfortran and cannot be built here.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PICNIC_REFERENCE:-/root/reference}"
SRC="$REF/src"
if [ ! -d "$SRC" ]; then echo "ref_build: $SRC not present (GPU box): keeping prebuilt oracle/_ref"; exit 0; fi
mkdir -p "$HERE/_ref"
FLAGS="-O2 -ffp-contract=off -fPIC -std=c++17 -DCH_SPACEDIM=2 -DCH_LANG_CC -w"
INC="-I$HERE/chombo_mock -I$SRC/particle_tools -I$SRC/species/pic -I$SRC/scattering -I$SRC/core"
g++ $FLAGS $INC -shared -o "$HERE/_ref/libpicnic_ref.so" \
  "$HERE/ref_driver.cpp" \
  "$SRC/species/pic/PicSpeciesUtils.cpp" \
  "$SRC/scattering/ScatteringUtils.cpp" \
  "$SRC/particle_tools/JustinsParticle.cpp" \
  "$SRC/particle_tools/BinItem.cpp"
echo "built $HERE/_ref/libpicnic_ref.so"
# the same sources as the reference compiles them with -DRELATIVISTIC_PARTICLES (Boris with gamma / Higuera-Cary,
# getImplicitGamma): pins the relativistic branch of the oracle
g++ $FLAGS -DRELATIVISTIC_PARTICLES $INC -shared -o "$HERE/_ref/libpicnic_ref_rel.so" \
  "$HERE/ref_driver.cpp" \
  "$SRC/species/pic/PicSpeciesUtils.cpp" \
  "$SRC/scattering/ScatteringUtils.cpp" \
  "$SRC/particle_tools/JustinsParticle.cpp" \
  "$SRC/particle_tools/BinItem.cpp"
echo "built $HERE/_ref/libpicnic_ref_rel.so"
# the reference's C++ gathers (MeshInterpI.H:1013-1850, the NEW_EM_INTERP_METHOD path) in 1D and 2D: pins the gather
# restatement of the oracle (tests/test_ref_pin_gather.py)
for D in 1 2; do
  g++ -O2 -ffp-contract=off -fPIC -std=c++17 -DCH_SPACEDIM=$D -DCH_LANG_CC -w -DREFMI_DEFINE_REALVECT_ZERO \
    -I"$HERE/chombo_mock" -I"$SRC/particle_tools" -I"$SRC/species/pic" -I"$SRC/core" \
    -shared -o "$HERE/_ref/libpicnic_ref_mi${D}d.so" \
    "$HERE/ref_meshinterp.cpp" "$SRC/particle_tools/JustinsParticle.cpp" "$SRC/particle_tools/BinItem.cpp"
  echo "built $HERE/_ref/libpicnic_ref_mi${D}d.so"
done
