"""ctypes binding of oracle/_ref/libpu_ref.so -- the reference's own CPU simulation path.

TEST INFRASTRUCTURE ONLY.  The library is built by oracle/build_ref.sh from the sources under
/root/reference (src/Sim/BruteForceCPU.cpp, BarnesHut.cpp, Octree.cpp, the seeders) and is the
"kind: reference" oracle and CPU baseline.  It may be imported from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs, never from the
product package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpu_ref.so")

# Layout of the reference's `struct Particle` (src/Render/Misc/Particle.hpp:8-18) under g++:
# Position float3 @0, Colour float4 @12, OriginalColour float4 @28, pad @44,
# Velocity double3 @48, Forces double3 @72, Mass double @96; sizeof == 104.
PARTICLE_DTYPE = np.dtype(
    {
        "names": ["Position", "Colour", "OriginalColour", "Velocity", "Forces", "Mass"],
        "formats": [("<f4", 3), ("<f4", 4), ("<f4", 4), ("<f8", 3), ("<f8", 3), "<f8"],
        "offsets": [0, 12, 28, 48, 72, 96],
        "itemsize": 104,
    }
)

SEED_RANDOM, SEED_GALAXY, SEED_STARSYSTEM = 0, 1, 2

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libpu_ref.so missing: run oracle/build_ref.sh where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp, sz, i64p, dp, ip = C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ref_sizeof_particle.restype = C.c_int
        L.ref_particle_offsets.argtypes = [ip]
        L.ref_hardware_workers.restype = C.c_int
        L.ref_seed.argtypes = [vp, sz, C.c_int, C.c_uint64, C.c_float]
        fp = C.POINTER(C.c_float)
        L.ref_sizeof_lwparticle.restype = C.c_int
        L.ref_seed_ex.argtypes = [vp, sz, C.c_int, C.c_uint64, C.c_float, fp]
        L.ref_seed_lw.argtypes = [vp, sz, C.c_int, C.c_uint64, C.c_float, fp]
        L.ref_closest_particle.argtypes = [vp, sz, fp]
        L.ref_closest_particle.restype = C.c_uint64
        L.ref_bruteforce_forces.argtypes = [vp, sz, i64p, sz, dp]
        L.ref_bruteforce_run.argtypes = [vp, sz, C.c_float, C.c_int, C.c_int, ip]
        L.ref_bruteforce_run.restype = C.c_double
        L.ref_bruteforce_block.argtypes = [vp, sz, sz, sz, C.c_int, ip, dp]
        L.ref_bruteforce_block.restype = C.c_double
        L.ref_barneshut_run.argtypes = [vp, sz, C.c_float, C.c_int, C.c_double, C.c_int, ip]
        L.ref_barneshut_run.restype = C.c_double
        L.ref_barneshut_forces.argtypes = [vp, sz, C.c_double, i64p, sz, dp, dp]
        L.ref_barneshut_forces.restype = C.c_double
        L.ref_octree_paths.argtypes = [vp, sz, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), i64p]
        L.ref_octree_cell.argtypes = [vp, sz, C.c_int, C.c_uint64, dp, C.POINTER(C.c_float),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.ref_octree_cell.restype = C.c_int
        L.ref_octree_leaf_cubes.argtypes = [vp, sz, fp, i64p]
        L.ref_octree_leaf_cubes.restype = C.c_size_t
        L.ref_barneshut_work.argtypes = [vp, sz, C.c_double, i64p, sz, i64p]
        L.ref_report_theta.argtypes = [C.c_float]
        L.ref_get_theta.restype = C.c_double
        assert L.ref_sizeof_particle() == PARTICLE_DTYPE.itemsize
        _lib = L
    return _lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _dbl(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def seed(n, kind=SEED_GALAXY, seed=42, scale=1.0):
    """Particles produced by the reference's CreateParticleSeeder(...)->Seed(seed)."""
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    lib().ref_seed(_vp(p), n, kind, seed, scale)
    return p


# The renderer's record (src/Render/Misc/Particle.hpp:20-25): Position float3, Colour float4, Scale float.
LWPARTICLE_DTYPE = np.dtype(
    {"names": ["Position", "Colour", "Scale"], "formats": [("<f4", 3), ("<f4", 4), "<f4"], "offsets": [0, 12, 28],
     "itemsize": 32}
)


def _rgb(colours):
    if colours is None:
        return None
    a = (C.c_float * 6)(*[float(x) for x in np.asarray(colours).reshape(6)])
    return a


def seed_ex(n, kind=SEED_GALAXY, seed=42, scale=1.0, colours=None, lw=False):
    """CreateParticleSeeder<T>(v, kind, scale) [+ Set{Red,Green,Blue}Dist] -> Seed(seed) for
    T = Particle or (lw=True) T = LWParticle.  colours = ((r_lo, r_hi), (g_lo, g_hi), (b_lo, b_hi))."""
    if lw:
        assert lib().ref_sizeof_lwparticle() == LWPARTICLE_DTYPE.itemsize
        p = np.zeros(n, dtype=LWPARTICLE_DTYPE)
        lib().ref_seed_lw(_vp(p), n, kind, seed, scale, _rgb(colours))
    else:
        p = np.zeros(n, dtype=PARTICLE_DTYPE)
        lib().ref_seed_ex(_vp(p), n, kind, seed, scale, _rgb(colours))
    return p


def closest_particle(p, pos):
    """Maths::ClosestParticle(pos, particles, &id) -> id."""
    q = (C.c_float * 3)(*[float(x) for x in pos])
    return int(lib().ref_closest_particle(_vp(p), len(p), q))


def bruteforce_forces(p, targets):
    """Forces (not accelerations) on the listed targets from BruteForceCPU::Exec."""
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.zeros((len(targets), 3), dtype=np.float64)
    lib().ref_bruteforce_forces(_vp(p), len(p), _i64(targets), len(targets), _dbl(out))
    return out


def bruteforce_run(p, dt, steps, workers=0):
    """steps x BruteForceCPU::Update(dt); returns (particles, seconds, workers used)."""
    q = p.copy()
    w = C.c_int(0)
    secs = lib().ref_bruteforce_run(_vp(q), len(q), dt, steps, workers, C.byref(w))
    return q, secs, w.value


def bruteforce_block(p, first, count, workers=0, want_forces=False):
    w = C.c_int(0)
    out = np.zeros((count, 3), dtype=np.float64) if want_forces else None
    secs = lib().ref_bruteforce_block(_vp(p), len(p), first, count, workers, C.byref(w),
                                      _dbl(out) if out is not None else None)
    return secs, w.value, out


def barneshut_run(p, dt, steps, theta=0.5, workers=0):
    q = p.copy()
    w = C.c_int(0)
    secs = lib().ref_barneshut_run(_vp(q), len(q), dt, steps, theta, workers, C.byref(w))
    return q, secs, w.value


def barneshut_forces(p, targets, theta=0.5):
    """Forces on targets from Octree::CalculateForce on a tree built as BarnesHut::Update does.
    Returns (forces, build seconds, traversal seconds)."""
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.zeros((len(targets), 3), dtype=np.float64)
    ev = C.c_double(0)
    build = lib().ref_barneshut_forces(_vp(p), len(p), theta, _i64(targets), len(targets), _dbl(out), C.byref(ev))
    return out, build, ev.value


def octree_paths(p):
    """(leaf depth per body, packed child-index path per body, stats dict)."""
    n = len(p)
    depth = np.zeros(n, dtype=np.int32)
    path = np.zeros(n, dtype=np.uint64)
    stats = np.zeros(4, dtype=np.int64)
    lib().ref_octree_paths(_vp(p), n, depth.ctypes.data_as(C.POINTER(C.c_int32)),
                           path.ctypes.data_as(C.POINTER(C.c_uint64)), _i64(stats))
    return depth, path, dict(nodes=int(stats[0]), internal=int(stats[1]), max_depth=int(stats[2]),
                             root_count=int(stats[3]))


def octree_leaf_cubes(p):
    """What Octree::RenderDebug draws: (cubes[k] = {pos.xyz, size}, body[k]) for every occupied leaf, in the
    reference's recursion order (children 0..7 = Morton order)."""
    n = len(p)
    out = np.zeros((n, 4), dtype=np.float32)
    body = np.zeros(n, dtype=np.int64)
    k = lib().ref_octree_leaf_cubes(_vp(p), n, out.ctypes.data_as(C.POINTER(C.c_float)), _i64(body))
    return out[:k], body[:k]


def octree_cell(p, depth, path):
    """(mass, com[3], population, width) of the cell at (depth, path), or None if the path leaves the tree."""
    m = C.c_double(0)
    com = (C.c_float * 3)()
    cnt = C.c_int32(0)
    wid = C.c_float(0)
    rc = lib().ref_octree_cell(_vp(p), len(p), depth, int(path), C.byref(m), com, C.byref(cnt), C.byref(wid))
    if rc != 0:
        return None
    return m.value, np.array(list(com), dtype=np.float32), cnt.value, wid.value


def barneshut_work(p, targets, theta=0.5):
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    out = np.zeros(3, dtype=np.int64)
    lib().ref_barneshut_work(_vp(p), len(p), theta, _i64(targets), len(targets), _i64(out))
    return dict(cell_evals=int(out[0]), leaf_evals=int(out[1]), visits=int(out[2]))
