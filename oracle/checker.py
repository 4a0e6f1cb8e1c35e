"""The checker used by tests, smoke() and bench.py: the reference's own CPU path (oracle/_ref,
kind "reference") when its prebuilt library is present, else the C restatement (kind "port").

TEST INFRASTRUCTURE ONLY.  Nothing under procedural-universe_b200/ imports this.
"""
import numpy as np

from . import port, ref


def kind():
    return "reference" if ref.available() else "port"


def allpairs_accel(p, targets=None):
    """Accelerations Forces/Mass of BruteForceCPU::Exec (BruteForceCPU.cpp:25-43) for the targets."""
    targets = np.arange(len(p)) if targets is None else np.asarray(targets)
    if ref.available():
        f = ref.bruteforce_forces(p, targets)
    else:
        f = np.concatenate([port.allpairs_forces(p, int(t), 1) for t in targets])
    return f / p["Mass"][targets][:, None]


def allpairs_run(p, dt, steps):
    if ref.available():
        return ref.bruteforce_run(p, dt, steps)[0]
    return port.allpairs_run(p, dt, steps)


def barneshut_accel(p, theta=0.5, targets=None):
    targets = np.arange(len(p)) if targets is None else np.asarray(targets)
    if ref.available():
        f = ref.barneshut_forces(p, targets, theta)[0]
    else:
        f = port.barneshut_forces(p, targets, theta)
    return f / p["Mass"][targets][:, None]


def barneshut_run(p, dt, steps, theta=0.5):
    if ref.available():
        return ref.barneshut_run(p, dt, steps, theta)[0]
    return port.barneshut_run(p, dt, steps, theta)


COLLISION = dict(separation=2000.0, approach_speed=2e16)     # bench.py's two-galaxy scene (nb_seed_collision_host)


def seed_scene(n, scene="galaxy", seed=42):
    """A workload's bodies through the REFERENCE's own GalaxySeeder (oracle/_ref): one galaxy, or the
    two-galaxy collision composed exactly as nb_seed_collision_host composes it (seeds s and s+1, n/2
    bodies each, centres -/+ separation/2 along x, approaching).  Needs oracle/_ref."""
    if scene != "collision":
        return ref.seed(n, ref.SEED_GALAXY, seed, 1.0)
    half = n // 2
    p = np.zeros(n, dtype=ref.PARTICLE_DTYPE)
    p[:half] = ref.seed(half, ref.SEED_GALAXY, seed, 1.0)
    p[half:] = ref.seed(n - half, ref.SEED_GALAXY, seed + 1, 1.0)
    sign = np.where(np.arange(n) < half, -1.0, 1.0)
    p["Position"][:, 0] += (sign * 0.5 * COLLISION["separation"]).astype(np.float32)
    p["Velocity"][:, 0] -= sign * COLLISION["approach_speed"]
    return p
