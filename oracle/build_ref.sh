#!/usr/bin/env bash
# Build the UNMODIFIED reference (sources read where they lie under $TPT_REFERENCE,
# default /root/reference) plus oracle/ref_harness.cpp into oracle/_ref/libtptref.so.
# Nothing is copied: the two files that need a compiler-compat edit are piped through
# sed straight into g++.
#   -fpermissive         MSVC-isms at BDPT.hpp:113 and BDPT.cpp:285
#   BDPT.cpp:141         `auto& lastVertex` binds a temporary: hard error in GCC -> `auto`
#   BDPT.cpp:41          `inline` dropped from GenerateCameraPath so the harness can link to it
#   -include stdlib.h    makes unqualified abs(float) the floating overload
#   PathTracer.cpp twice as shipped, and with the stray `break;` of line 109 deleted and
#                        the function renamed PathTraceFull (the README PT images)
# Plain -O3, no -march: no FMA contraction, this is the parity oracle.
set -euo pipefail
REF="${TPT_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_ref"
if [ ! -f "$REF/BDPT.cpp" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt $OUT (if any)" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
CXX="${CXX:-g++}"
FLAGS="-std=c++17 -O3 -DNDEBUG -fpermissive -w -fPIC -include stdlib.h -I$REF"
pids=()
for f in Vector BVH Triangle Sphere Scene Material global Random SampleHelperFunctions SceneRenderingHelper Renderer PathTracer; do
  $CXX $FLAGS -c "$REF/$f.cpp" -o "$OUT/obj/$f.o" & pids+=($!)
done
sed -e '141s/auto& lastVertex/auto lastVertex/' -e '41s/^inline void BDPTPath::GenerateCameraPath/void BDPTPath::GenerateCameraPath/' "$REF/BDPT.cpp" | $CXX $FLAGS -x c++ -c - -o "$OUT/obj/BDPT.o" & pids+=($!)
sed -e '109d' -e 's/^Vector3f PathTrace(/Vector3f PathTraceFull(/' "$REF/PathTracer.cpp" \
  | $CXX $FLAGS -x c++ -c - -o "$OUT/obj/PathTracerFull.o" & pids+=($!)
$CXX $FLAGS -I"$ROOT/include" -I"$ROOT/toypathtracer-games101-assignment7_b200/host" \
  -c "$HERE/ref_harness.cpp" -o "$OUT/obj/ref_harness.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
$CXX -shared -o "$OUT/libtptref.so" "$OUT"/obj/*.o -lpthread
echo "built $OUT/libtptref.so"
