"""Regenerates tests/golden/*.npz from the reference's own CPU path (oracle/_ref/libpu_ref.so,
built by oracle/build_ref.sh from /root/reference/src/Sim).  Run where /root/reference exists:

    python tests/golden/make_golden.py

The vectors pin (a) the oracle restatement oracle/nbody_port.c, (b) the product's host seeder and
(c) the CUDA path, on machines where the reference itself is not available.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # 1. galaxy seeds (IParticleSeeder::Seed through CreateParticleSeeder)
    seeds = {}
    for n, seed, scale in [(256, 42, 1.0), (1000, 7, 2.5), (4096, 42, 1.0)]:
        p = ref.seed(n, ref.SEED_GALAXY, seed, scale)
        seeds[f"n{n}_s{seed}"] = p.view(np.uint8).reshape(n, 104)
    np.savez_compressed(os.path.join(HERE, "galaxy_seeds.npz"), **seeds)

    # 2. all-pairs: forces at step 0 and state after 10 x Update(0.01), N = 256
    p = ref.seed(256, ref.SEED_GALAXY, 42, 1.0)
    f0 = ref.bruteforce_forces(p, np.arange(256))
    q, _, w = ref.bruteforce_run(p, 0.01, 10, workers=4)
    np.savez_compressed(os.path.join(HERE, "allpairs_n256.npz"), forces0=f0,
                        state10=q.view(np.uint8).reshape(256, 104), dt=np.float32(0.01), steps=10, workers=w)

    # 3. config 1 of BASELINE.json: N = 4096, dt = 0.01, 100 steps (positions/velocities only)
    p = ref.seed(4096, ref.SEED_GALAXY, 42, 1.0)
    f0 = ref.bruteforce_forces(p, np.arange(0, 4096, 16))
    q, _, w = ref.bruteforce_run(p, 0.01, 100, workers=4)
    np.savez_compressed(os.path.join(HERE, "allpairs_n4096_100steps.npz"), targets=np.arange(0, 4096, 16),
                        forces0=f0, pos=q["Position"].copy(), vel=q["Velocity"].copy(), workers=w)

    # 4. Barnes-Hut theta = 0.5: forces, octree paths, work counters, 5 steps at dt = 0.02/60
    p = ref.seed(1024, ref.SEED_GALAXY, 42, 1.0)
    t = np.arange(1024)
    fb, _, _ = ref.barneshut_forces(p, t, 0.5)
    depth, path, stats = ref.octree_paths(p)
    work = ref.barneshut_work(p, t, 0.5)
    q, _, w = ref.barneshut_run(p, np.float32(0.02 / 60), 5, 0.5, workers=4)
    np.savez_compressed(os.path.join(HERE, "barneshut_n1024.npz"), forces=fb, leaf_depth=depth, path=path,
                        max_depth=stats["max_depth"], nodes=stats["nodes"],
                        work=np.array([work["cell_evals"], work["leaf_evals"], work["visits"]]),
                        state5=q.view(np.uint8).reshape(1024, 104), dt=np.float32(0.02 / 60))
    # 5. energy drift of the reference Barnes-Hut path on the two-galaxy collision scene (config 5 at
    #    a CPU-feasible size): N = 4096, seeds 42/43, separation 2000, approach 2e16, theta 0.5,
    #    dt = 0.02/60, 1000 steps; E = sum 1/2 m v^2 + Scale * sum U(r) evaluated exactly (all pairs).
    from oracle import port
    n = 4096
    a = ref.seed(n // 2, ref.SEED_GALAXY, 42, 1.0)
    b = ref.seed(n // 2, ref.SEED_GALAXY, 43, 1.0)
    p = np.concatenate([a, b]).astype(ref.PARTICLE_DTYPE)
    scene = np.zeros(n, dtype=ref.PARTICLE_DTYPE)
    for name in scene.dtype.names:
        scene[name] = np.concatenate([a[name], b[name]])
    sign = np.where(np.arange(n) < n // 2, -1.0, 1.0)
    scene["Position"][:, 0] += (sign * 0.5 * 2000.0).astype(np.float32)
    scene["Velocity"][:, 0] -= sign * 2e16
    dt = np.float32(0.02 / 60)
    energies = [port.energy(scene)]
    q = scene
    for _ in range(10):
        q, _, _ = ref.barneshut_run(q, dt, 100, 0.5, workers=4)
        energies.append(port.energy(q))
    e = np.array(energies)
    tot = e.sum(axis=1)
    drift = np.abs(tot - tot[0]) / abs(tot[0])
    print("reference BH energy drift per 100 steps:", drift)
    np.savez_compressed(os.path.join(HERE, "energy_drift_n4096.npz"), scene=scene.view(np.uint8).reshape(n, 104),
                        energies=e, drift=drift, dt=dt, theta=0.5, separation=2000.0, approach=2e16)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
