#!/usr/bin/env python
"""The reference's in-app benchmark (SimulationState::RunBenchmark, SimulationState.cpp:334-362: per sim type,
Init then 10 x Update(1.0f), ms per frame) run through INBodySim for the reference's own CPU sims and for the B200
adapter, at the particle counts the reference's UI allows (UI.cpp:75 caps N at 50 000).  Needs oracle/_ref (built
where /root/reference exists; travels to the GPU box) and a B200.

    python tools/benchmark_protocol.py [n ...]        -> one table per n
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("procedural-universe_b200")
from oracle import ref  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libb200_adapter_test.so"))
lib.adapter_benchmark.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
lib.adapter_benchmark.restype = C.c_double
hw = ref.lib().ref_hardware_workers()

PROTOCOLS = ((1.0, "RunBenchmark's Update(1.0f): every body leaves the octree's root cube in the first frame"),
             (0.02 / 60, "Update(dt * SimSpeed) of an interactive frame, dt = 0.02 / 60"))
SIMS = (("BruteForceCPU (reference)", 0, 0), ("BarnesHut (reference, theta 2.0)", 0, 2),
        ("B200Sim all-pairs (replaces BruteForceGPU)", 1, 1), ("B200Sim Barnes-Hut (theta 2.0)", 1, 2))

for n in [int(x) for x in sys.argv[1:]] or [1000, 4000, 50000]:
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    workers = max(w for w in range(1, hw + 1) if n % w == 0)        # the reference mis-indexes the remainder when W does not divide N
    for dt, what in PROTOCOLS:
        print(f"N = {n}, {what}  (reference pool: {workers} workers of {hw}; theta = the reference's default 2.0; "
              f"ms per Update through INBodySim, host array in and out)")
        for name, impl, kind in SIMS:
            frames = 2 if (impl == 0 and kind == 0 and n > 4000) else 10
            q = p.copy()
            lib.adapter_benchmark(q.ctypes.data, n, impl, kind, 2, workers, 2.0, dt)      # warm-up: contexts, pools, allocations
            q = p.copy()
            ms = lib.adapter_benchmark(q.ctypes.data, n, impl, kind, frames, workers, 2.0, dt)
            print(f"    {name:46s} {ms:12.3f} ms/frame   ({frames} frames)", flush=True)
