#!/usr/bin/env python
"""Energy drift of the REFERENCE Barnes-Hut path (oracle/_ref) on the two-galaxy collision scene at
N = 65536 -- the largest size its CPU path finishes in minutes (0.55 s/step on 4 workers here).
Same scene definition, theta, dt and estimator as energy_drift_n4096.npz (make_golden.py, item 5),
16x the bodies: total mass and therefore the drift are much larger, which makes it the sharper
comparison for the GPU path.  Writes tests/golden/energy_drift_n65536.npz (energies + drift only;
the scene is regenerated from the seeds by the bit-exact host seeder)."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402

pkg = importlib.import_module("procedural-universe_b200")

n, every, chunks = 65536, 250, 4
scene = pkg.seed_collision_host(n, 42, 1.0, separation=2000.0, approach_speed=2e16)
a = ref.seed(n // 2, ref.SEED_GALAXY, 42, 1.0)
assert np.array_equal(a["Position"][:, 1], scene["Position"][: n // 2, 1])      # same bodies as the reference seeder's
dt = np.float32(0.02 / 60)
energies = [port.energy(scene)]
q = scene
for k in range(chunks):
    q, _, _ = ref.barneshut_run(q, dt, every, 0.5, workers=4)
    energies.append(port.energy(q))
    print(k, energies[-1], flush=True)
e = np.array(energies)
tot = e.sum(axis=1)
drift = np.abs(tot - tot[0]) / abs(tot[0])
print("reference BH energy drift every", every, "steps:", drift)
np.savez_compressed(os.path.join(HERE, "energy_drift_n65536.npz"), energies=e, drift=drift, dt=dt, theta=0.5, n=n, every=every,
                    separation=2000.0, approach=2e16, final_pos_sample=q["Position"][::1024].copy(),
                    final_vel_sample=q["Velocity"][::1024].copy())
