#!/bin/bash
# Build UNMODIFIED reference programs against the drop-in headers of this repository.
#   tools/build_dropin.sh [reference root] [program ...]      e.g. tools/build_dropin.sh /root/reference test/cavityflow3D.cpp
# A scratch overlay tree is assembled under build/dropin/ from symbolic links only (no reference source is copied into the
# repository): <overlay>/src = this repository's panslbm2_b200/src plus links to the reference's untouched host-side utilities
# (MMA optimiser, VTK writers), <overlay>/{test,production}/*.cpp = links to the reference programs.  Their
# `#include "../src/..."` lines then resolve to the drop-in headers.  Binaries land in build/dropin/bin/ (git-ignored; they
# travel to the GPU box with gpurun).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${1:-/root/reference}
shift || true
OV=$ROOT/build/dropin
rm -rf "$OV/src" "$OV/test" "$OV/production"
mkdir -p "$OV/src/utility" "$OV/test" "$OV/production" "$OV/bin"
for d in b200 particle equation; do ln -s "$ROOT/panslbm2_b200/src/$d" "$OV/src/$d"; done
for f in "$ROOT"/panslbm2_b200/src/utility/*.h; do ln -s "$f" "$OV/src/utility/$(basename "$f")"; done
for f in "$REF"/src/utility/*.h; do [ -e "$OV/src/utility/$(basename "$f")" ] || ln -s "$f" "$OV/src/utility/$(basename "$f")"; done
PROGS=("$@")
[ ${#PROGS[@]} -eq 0 ] && PROGS=(test/cavityflow3D.cpp test/cavityflow.cpp test/d2q9.cpp test/d3q15.cpp test/nsadncsens.cpp test/nssens3D.cpp test/nssens.cpp test/nsadsens.cpp test/naturalconvection.cpp production/heatsink.cpp production/heatsink3D.cpp production/ncpump.cpp production/nsopt.cpp \
  test/heavisidefilter.cpp production/heatsink3D_transient.cpp production/heatsink_transient.cpp production/ncpump_periodic.cpp)
unset CC CXX
for p in "${PROGS[@]}"; do
  ln -sf "$REF/$p" "$OV/$p"
  out="$OV/bin/$(basename "${p%.cpp}")"
  if g++ -O2 -mavx -w -I"$ROOT/include" -I"$ROOT/panslbm2_b200/src/mpi" "$OV/$p" -o "$out" -L"$ROOT/panslbm2_b200" -lpanslbm_b200 -Wl,-rpath,'$ORIGIN/../../../panslbm2_b200' 2> "$out.log"; then
    echo "built $out"; rm -f "$out.log"
  else
    echo "FAILED $p (see $out.log)"; head -20 "$out.log"
  fi
done
