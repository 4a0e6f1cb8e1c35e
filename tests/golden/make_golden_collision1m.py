#!/usr/bin/env python
"""The REFERENCE Barnes-Hut path (oracle/_ref: BarnesHut.cpp + Octree.cpp as they lie under
/root/reference) on the two-galaxy collision scene of BASELINE.json configs[4] at N = 2^20 for 100
steps (theta 0.5, dt 0.02/60) -- about 40 minutes of its CPU path on 4 pool workers.  Stores

  * the sampled-energy estimator (oracle/port.py energy_sampled == nb_energy_sampled, every 256th
    body x all sources) at steps 0, 25, 50, 75, 100 and the drift |E - E0| / |E0|,
  * position and velocity of every 256th body after 100 steps (trajectory parity at 1 M bodies),
  * the sha256 of the seeded array (the GPU test re-seeds with the product's host seeder).

    python tests/golden/make_golden_collision1m.py        -> tests/golden/collision_n1048576_100steps.npz
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import port, ref  # noqa: E402
from make_golden_bh16m import seed_scene  # noqa: E402

n, every, chunks, stride = 1 << 20, 25, 4, 256
scene = seed_scene(n, "collision")
sha = hashlib.sha256(scene.view(np.uint8)).hexdigest()
dt = np.float32(0.02 / 60)
t0 = time.time()
energies = [port.energy_sampled(scene, stride)[:2]]
print("E0", energies[0], f"{time.time() - t0:.0f} s", flush=True)
q = scene
for k in range(chunks):
    q, secs, w = ref.barneshut_run(q, dt, every, 0.5, workers=4)
    energies.append(port.energy_sampled(q, stride)[:2])
    print(k, energies[-1], f"{secs:.0f} s on {w} workers", flush=True)
e = np.array(energies)
tot = e.sum(axis=1)
drift = np.abs(tot - tot[0]) / abs(tot[0])
print("reference BH energy drift every", every, "steps:", drift)
np.savez_compressed(os.path.join(HERE, "collision_n1048576_100steps.npz"), energies=e, drift=drift, dt=dt, theta=0.5, n=n,
                    every=every, stride=stride, separation=2000.0, approach=2e16, sha256=np.array(sha),
                    final_pos_sample=q["Position"][::stride].copy(), final_vel_sample=q["Velocity"][::stride].copy())
