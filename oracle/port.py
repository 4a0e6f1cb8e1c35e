"""ctypes binding of oracle/_build/libnbody_port.so (the plain-C restatement, nbody_port.c).

TEST INFRASTRUCTURE ONLY -- see the header of nbody_port.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libnbody_port.so")
MORTON_OUTSIDE = np.uint64(0xFFFFFFFFFFFFFFFF)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "nbody_port.c")):
            subprocess.run(["make", "-C", _HERE, "_build/libnbody_port.so"], check=True, capture_output=True)
        L = C.CDLL(LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.port_allpairs_forces.argtypes = [vp, sz, sz, sz, vp]
        L.port_allpairs_run.argtypes = [vp, sz, C.c_float, C.c_int]
        L.port_barneshut_forces.argtypes = [vp, sz, C.c_double, vp, sz, vp, vp]
        L.port_barneshut_run.argtypes = [vp, sz, C.c_float, C.c_int, C.c_double]
        L.port_octree_paths.argtypes = [vp, sz, vp, vp, vp]
        L.port_morton.argtypes = [vp, sz, vp]
        L.port_morton_one.argtypes = [C.c_float, C.c_float, C.c_float]
        L.port_morton_one.restype = C.c_uint64
        L.port_morton_sorted.argtypes = [vp, sz, vp, vp]
        L.port_morton_sorted.restype = sz
        L.port_karras.argtypes = [vp, sz, vp, vp, vp]
        L.port_energy.argtypes = [vp, sz, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def allpairs_forces(p, first=0, count=None):
    count = len(p) - first if count is None else count
    out = np.zeros((count, 3), dtype=np.float64)
    lib().port_allpairs_forces(p.ctypes.data, len(p), first, count, out.ctypes.data)
    return out


def allpairs_run(p, dt, steps):
    q = p.copy()
    lib().port_allpairs_run(q.ctypes.data, len(q), dt, steps)
    return q


def barneshut_forces(p, targets, theta=0.5, want_counters=False):
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.zeros((len(targets), 3), dtype=np.float64)
    cnt = np.zeros(3, dtype=np.int64)
    lib().port_barneshut_forces(p.ctypes.data, len(p), theta, targets.ctypes.data, len(targets), out.ctypes.data, cnt.ctypes.data)
    if want_counters:
        return out, dict(cell_evals=int(cnt[0]), leaf_evals=int(cnt[1]), visits=int(cnt[2]))
    return out


def barneshut_run(p, dt, steps, theta=0.5):
    q = p.copy()
    lib().port_barneshut_run(q.ctypes.data, len(q), dt, steps, theta)
    return q


def octree_paths(p):
    n = len(p)
    depth = np.zeros(n, dtype=np.int32)
    path = np.zeros(n, dtype=np.uint64)
    stats = np.zeros(4, dtype=np.int64)
    lib().port_octree_paths(p.ctypes.data, n, depth.ctypes.data, path.ctypes.data, stats.ctypes.data)
    return depth, path, dict(max_depth=int(stats[2]), root_count=int(stats[3]))


def morton(p):
    codes = np.zeros(len(p), dtype=np.uint64)
    lib().port_morton(p.ctypes.data, len(p), codes.ctypes.data)
    return codes


def morton_sorted(p):
    codes = np.zeros(len(p), dtype=np.uint64)
    order = np.zeros(len(p), dtype=np.uint32)
    m = lib().port_morton_sorted(p.ctypes.data, len(p), codes.ctypes.data, order.ctypes.data)
    return codes[:m].copy(), order[:m].copy()


def karras(sorted_codes):
    k = np.ascontiguousarray(sorted_codes, dtype=np.uint64)
    m = len(k)
    left = np.zeros(max(m - 1, 0), dtype=np.int32)
    right = np.zeros(max(m - 1, 0), dtype=np.int32)
    prefix = np.zeros(max(m - 1, 0), dtype=np.int32)
    if m >= 2:
        lib().port_karras(k.ctypes.data, m, left.ctypes.data, right.ctypes.data, prefix.ctypes.data)
    return left, right, prefix


def energy(p):
    ke, pe = C.c_double(), C.c_double()
    lib().port_energy(p.ctypes.data, len(p), C.byref(ke), C.byref(pe))
    return ke.value, pe.value


def energy_sampled(p, stride, G=6.674e-11, S=10.0, scale=20 * 1.15e12):
    """The estimator of nb_energy_sampled (csrc/energy.cu) restated in numpy: kinetic energy exactly,
    potential from the bodies whose index is a multiple of `stride`, each against ALL sources, times `stride`:
    E = sum 1/2 m v^2 + scale * sum_{i<j} U(r),  U = -(G ma mb / sqrt(S)) atan(sqrt(S) / r).
    Returns (kinetic, potential estimate, samples)."""
    pos = p["Position"].astype(np.float64)
    m = p["Mass"]
    ke = float(np.sum(0.5 * m * np.sum(p["Velocity"] ** 2, axis=1)))
    w = (G * m).astype(np.float32).astype(np.float64)      # the device keeps G m_j in fp32
    sq = np.sqrt(S)
    pe = 0.0
    idx = np.arange(0, len(p), stride)
    for i in idx:
        d = pos - pos[i]
        r = np.sqrt(np.einsum("ij,ij->i", d, d))
        with np.errstate(divide="ignore"):
            a = np.where(r > 0.0, np.arctan(sq / r), np.pi / 2)
        a[i] = 0.0
        pe += 0.5 * m[i] / sq * float(-(w * a).sum())
    return ke, pe * stride * scale, len(idx)


def recentre(p):
    """InitParticlesFromFile's recentring (reference SimulationState.cpp:252-270) restated in numpy:
    TotalMass accumulates in `long double`, the weighted position sums in double, both in index
    order (cumsum is a sequential scan); CentreOfMass /= (double)TotalMass; Position -= (float3)centre."""
    q = p.copy()
    if len(q) == 0:
        return q
    m = q["Mass"]
    total = float(np.cumsum(m.astype(np.longdouble))[-1])
    pos = q["Position"].astype(np.float64)
    with np.errstate(all="ignore"):
        c = np.array([np.cumsum(pos[:, k] * m)[-1] for k in range(3)]) / total
        q["Position"] = q["Position"] - c.astype(np.float32)
    return q


def closest_particle(p, pos):
    """Maths::ClosestParticle (reference Core/Maths.hpp:62-85): fp32 ((dx*dx + dy*dy) + dz*dz), strict <
    from FLT_MAX in index order -> first index among equal minima, 0 when nothing qualifies."""
    q = np.asarray(pos, dtype=np.float32)
    with np.errstate(all="ignore"):
        d = q[None, :] - p["Position"]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    ok = d2 < np.finfo(np.float32).max
    if not ok.any():
        return 0
    return int(np.argmin(np.where(ok, d2, np.inf)))
