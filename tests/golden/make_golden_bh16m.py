"""Sampled-target Barnes-Hut goldens at the sizes BASELINE.json names for the tree code.

    python tests/golden/make_golden_bh16m.py [bh16m] [collision64m] [collision1m]

Run where /root/reference exists (oracle/_ref/libpu_ref.so = the reference's own BarnesHut.cpp /
Octree.cpp).  For each scene: seed through the REFERENCE's seeder, build the reference octree once
exactly as BarnesHut::Update does (BarnesHut.cpp:46-56: Add every particle, CalculateMass) and call
Octree::CalculateForce (Octree.cpp:107-145) at theta = 0.5 for a fixed sample of targets.  Host RAM:
~12 GB at N = 2^24, ~45 GB at N = 2^26 (136-byte nodes, ~3.9 N of them).

What is stored (small: a few KB per scene): target indices, the reference forces, the targets' own
records, and a checksum of the whole seeded array so the GPU test can prove it rebuilt the same bodies
from the same seeds without shipping gigabytes.  The direct sum (BruteForceCPU::Exec on the same targets)
is stored next to it: it separates "GPU tree differs from reference tree" from "both differ from exact".
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

SCENES = {
    # name: (n, kind, number of sampled targets)
    "bh16m": (1 << 24, "galaxy", 256),
    "collision64m": (1 << 26, "collision", 64),
    "collision1m": (1 << 20, "collision", 256),
}
COLLISION = dict(separation=2000.0, approach_speed=2e16)   # bench.py's scene (nb_seed_collision_host)


def seed_scene(n, kind):
    if kind == "galaxy":
        return ref.seed(n, ref.SEED_GALAXY, 42, 1.0)
    half = n // 2
    p = np.zeros(n, dtype=ref.PARTICLE_DTYPE)
    p[:half] = ref.seed(half, ref.SEED_GALAXY, 42, 1.0)
    p[half:] = ref.seed(n - half, ref.SEED_GALAXY, 43, 1.0)
    sign = np.where(np.arange(n) < half, -1.0, 1.0)
    p["Position"][:, 0] += (sign * 0.5 * COLLISION["separation"]).astype(np.float32)
    p["Velocity"][:, 0] -= sign * COLLISION["approach_speed"]
    return p


def sample_targets(n, k):
    """k targets spread over the index range (arm segments and disk bodies both get hit), fixed."""
    rng = np.random.RandomState(12345)
    t = np.unique(np.concatenate([np.arange(0, n, n // (k // 2)), rng.randint(0, n, size=k)]))[:k]
    return np.sort(t).astype(np.int64)


def checksum(p):
    return hashlib.sha256(p.view(np.uint8)).hexdigest()


def main():
    names = sys.argv[1:] or ["bh16m"]
    for name in names:
        n, kind, k = SCENES[name]
        t0 = time.time()
        p = seed_scene(n, kind)
        print(f"{name}: seeded {n} bodies in {time.time() - t0:.1f} s", flush=True)
        targets = sample_targets(n, k)
        f, build, walk = ref.barneshut_forces(p, targets, 0.5)
        print(f"{name}: reference octree build {build:.1f} s, {len(targets)} x CalculateForce {walk:.2f} s", flush=True)
        work = ref.barneshut_work(p, targets[:16], 0.5) if n <= (1 << 24) else None
        t0 = time.time()
        direct = ref.bruteforce_forces(p, targets[: min(len(targets), 64)])
        print(f"{name}: direct sums {time.time() - t0:.1f} s", flush=True)
        finite = np.isfinite(f).all(axis=1)
        print(f"{name}: finite reference forces: {int(finite.sum())} of {len(targets)}")
        out = dict(n=n, theta=0.5, targets=targets, forces=f, records=p[targets].view(np.uint8).reshape(len(targets), 104),
                   direct_forces=direct, sha256=np.array(checksum(p)), build_s=build, walk_s=walk)
        if work is not None:
            out["work16"] = np.array([work["cell_evals"], work["leaf_evals"], work["visits"]])
        np.savez_compressed(os.path.join(HERE, f"bh_{name}_sampled.npz"), **out)
        rel = np.linalg.norm(f[: len(direct)] - direct, axis=1) / np.linalg.norm(direct, axis=1)
        print(f"{name}: reference tree vs reference direct sum: median {np.median(rel):.3e} max {rel.max():.3e}")
        del p


if __name__ == "__main__":
    main()
