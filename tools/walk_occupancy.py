#!/usr/bin/env python
"""Lane occupancy of the Barnes-Hut walk (nb_get_walk_occupancy): how many of a warp's 32 lanes are awake per node
visit, and how much of the walk's issue time the sparse visits take."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("procedural-universe_b200")
for n in [int(x) for x in sys.argv[1:]] or [1 << 20, 1 << 24]:
    sim = pkg.Sim(mode=pkg.MODE_BARNESHUT, theta=0.5)
    sim.seed_galaxy_device(n, 42, 1.0)
    st = sim.walk_stats()
    h = sim.walk_occupancy().astype(np.float64)
    it = h.sum()
    k = np.arange(33)
    print("n", n, st, "warp visits", it, "per warp", it / (n / 32))
    print(" lane slots awake %.3f" % ((h * k).sum() / (32 * it)))
    c = np.cumsum(h) / it
    lanes = np.cumsum(h * k) / (h * k).sum()
    for q in (1, 2, 4, 8, 16, 24, 31, 32):
        print("  <=%2d lanes awake: %.3f of the visits, %.3f of the awake lane-visits (%.1f per lane)" % (q, c[q], lanes[q], np.cumsum(h * k)[q] / n))
    sp = sim.walk_sparse_load().astype(np.float64)
    for i, q in enumerate((4, 8, 16)):
        total, mx, warps = sp[2 * i], sp[2 * i + 1], sp[6]
        print("  visits with <=%2d lanes awake: %.1f per lane on average, busiest lane of a warp %.1f on average (imbalance %.2f)"
              % (q, total / (32 * warps), mx / warps, mx / (total / 32)))
    sim.close()
