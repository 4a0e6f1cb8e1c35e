#!/usr/bin/env bash
# Builds oracle/_ref/libpu_ref.so: the reference's own CPU simulation path (src/Sim), compiled
# headless from the sources where they lie under /root/reference.  TEST INFRASTRUCTURE ONLY.
#
# Nothing from the reference is copied into the repository: the three files that need a patch to
# compile with g++ are copied to a scratch directory under the git-ignored oracle/_ref/, patched
# there, used for the build and deleted again.  The patches:
#   1. Sim/IParticleSeeder.hpp:8-10 includes the three seeder headers before it defines
#      IParticleSeeder (accepted by MSVC's lazy template parsing, rejected by g++): move the
#      includes below the class.
#   2. Sim/GalaxySeeder.cpp:53 names `Matrix` unqualified: add a using-declaration.
#   3. Core/ThreadPool.hpp:28 has room for 31 workers but spawns hardware_concurrency()-1:
#      widen the arrays so a many-core GPU host does not overflow them.
# BruteForceCPU.cpp, BarnesHut.cpp, Octree.cpp, Physics.hpp, Log.cpp, Event.cpp compile unmodified.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PU_REFERENCE:-/root/reference}/src"
OUT="$HERE/_ref"
if [ ! -d "$REF/Sim" ]; then
    echo "build_ref: $REF/Sim not found (reference absent); keeping any prebuilt $OUT/libpu_ref.so" >&2
    exit 0
fi
mkdir -p "$OUT"
SCRATCH="$(mktemp -d "$OUT/scratch.XXXXXX")"
trap 'rm -rf "$SCRATCH"' EXIT
mkdir -p "$SCRATCH/Sim" "$SCRATCH/Core"
for f in IParticleSeeder.hpp GalaxySeeder.hpp GalaxySeeder.cpp RandomSeeder.hpp RandomSeeder.cpp \
         StarSystemSeeder.hpp StarSystemSeeder.cpp; do
    cp "$REF/Sim/$f" "$SCRATCH/Sim/$f"
done
cp "$REF/Core/ThreadPool.hpp" "$SCRATCH/Core/ThreadPool.hpp"

python3 - "$SCRATCH" <<'PY'
import re, sys, pathlib
scratch = pathlib.Path(sys.argv[1])

p = scratch / "Sim" / "IParticleSeeder.hpp"
s = p.read_text()
incs = [l for l in s.splitlines() if re.match(r'#include "(Random|Galaxy|StarSystem)Seeder\.hpp"', l)]
assert len(incs) == 3, incs
for l in incs:
    s = s.replace(l + "\n", "", 1)
marker = "template <class T = Particle>"
assert marker in s
s = s.replace(marker, "\n".join(incs) + "\n\n" + marker, 1)
p.write_text(s)

p = scratch / "Sim" / "GalaxySeeder.cpp"
s = p.read_text()
anchor = "using DirectX::SimpleMath::Color;"
assert anchor in s
s = s.replace(anchor, anchor + "\nusing DirectX::SimpleMath::Matrix;", 1)
p.write_text(s)

p = scratch / "Core" / "ThreadPool.hpp"
s = p.read_text()
assert "MAX_WORKERS = 31" in s
s = s.replace("MAX_WORKERS = 31", "MAX_WORKERS = 1023", 1)
p.write_text(s)
PY

CXXFLAGS="-std=c++14 -O2 -ffp-contract=off -fno-access-control -pthread -fPIC -w"
g++ $CXXFLAGS -shared \
    -I"$HERE/ref_shim" -I"$SCRATCH" -I"$REF" -I"$REF/Sim" \
    "$HERE/ref_driver.cpp" \
    "$REF/Sim/BruteForceCPU.cpp" "$REF/Sim/BarnesHut.cpp" "$REF/Sim/Octree.cpp" \
    "$REF/Services/Log.cpp" "$REF/Core/Event.cpp" \
    -o "$OUT/libpu_ref.so"
echo "build_ref: wrote $OUT/libpu_ref.so"

# The product's C++ adapter (B200Sim : INBodySim) compiled against the same reference headers, plus
# a driver that runs reference sims and the adapter through the reference's own interface.
PKG="$HERE/../procedural-universe_b200"
if [ -f "$PKG/lib/libnbody_b200.so" ]; then
    g++ $CXXFLAGS -shared \
        -I"$HERE/ref_shim" -I"$SCRATCH" -I"$REF" -I"$REF/Sim" -I"$HERE/../include" -I"$PKG/host" \
        "$HERE/adapter_driver.cpp" "$PKG/host/B200Sim.cpp" \
        "$REF/Sim/BruteForceCPU.cpp" "$REF/Sim/BarnesHut.cpp" "$REF/Sim/Octree.cpp" \
        "$REF/Services/Log.cpp" "$REF/Core/Event.cpp" \
        -L"$PKG/lib" -lnbody_b200 -Wl,-rpath,'$ORIGIN/../../procedural-universe_b200/lib' \
        -o "$OUT/libb200_adapter_test.so"
    echo "build_ref: wrote $OUT/libb200_adapter_test.so"
else
    echo "build_ref: $PKG/lib/libnbody_b200.so not built yet, skipping the adapter test library" >&2
fi
