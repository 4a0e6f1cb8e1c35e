#!/usr/bin/env python
"""Times the device galaxy seeder (csrc/seed_device.cu) against the host seeder at a few sizes and checks equality.
Usage: python tools/seed_timing.py [n ...]      (run under ncu for the per-kernel list in profiles/)"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("procedural-universe_b200")

for n in [int(x) for x in sys.argv[1:]] or [1 << 20, 1 << 24]:
    sim = pkg.Sim(mode=pkg.MODE_ALLPAIRS)
    sim.seed_galaxy_device(1024, 1, 1.0)           # context, module load
    t0 = time.perf_counter()
    sim.seed_galaxy_device(n, 42, 1.0)
    t_dev = time.perf_counter() - t0
    pos, vel = sim.read_soa()
    sim.close()
    t0 = time.perf_counter()
    host = pkg.seed_galaxy_host(n, 42, 1.0)
    t_host = time.perf_counter() - t0
    same = bool(np.array_equal(pos, host["Position"]) and np.array_equal(vel, host["Velocity"]))
    print(f"n = {n}: device seeder (into the handle) {t_dev * 1e3:.1f} ms, host seeder {t_host * 1e3:.0f} ms, identical positions and velocities: {same}", flush=True)
