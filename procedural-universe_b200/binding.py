"""ctypes binding of include/nbody_b200.h (libnbody_b200.so).  No compute happens in Python."""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnbody_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nbody_b200.h")

MODE_ALLPAIRS = 0
MODE_BARNESHUT = 2

# reference src/Render/Misc/Particle.hpp:8-18
PARTICLE_DTYPE = np.dtype(
    {
        "names": ["Position", "Colour", "OriginalColour", "Velocity", "Forces", "Mass"],
        "formats": [("<f4", 3), ("<f4", 4), ("<f4", 4), ("<f8", 3), ("<f8", 3), "<f8"],
        "offsets": [0, 12, 28, 48, 72, 96],
        "itemsize": 104,
    }
)


class NBodyError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("mode", C.c_int32),
        ("theta", C.c_float),
        ("G", C.c_double),
        ("softening", C.c_double),
        ("position_scale", C.c_double),
        ("bounds", C.c_float),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("stream", C.c_void_p),
        ("source_splits", C.c_int32),
        ("kernel_variant", C.c_int32),
    ]


class SeedOptions(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("layout", C.c_int32),
        ("scale", C.c_float),
        ("red", C.c_float * 2),
        ("green", C.c_float * 2),
        ("blue", C.c_float * 2),
    ]


SEEDER_RANDOM, SEEDER_GALAXY, SEEDER_STARSYSTEM = 0, 1, 2
LAYOUT_PARTICLE, LAYOUT_LWPARTICLE = 0, 1
# The renderer's 32-byte record (reference src/Render/Misc/Particle.hpp:20-25).
LWPARTICLE_DTYPE = np.dtype(
    {"names": ["Position", "Colour", "Scale"], "formats": [("<f4", 3), ("<f4", 4), "<f4"], "offsets": [0, 12, 28],
     "itemsize": 32}
)

_lib = None


def declared_symbols():
    """Every function name include/nbody_b200.h declares."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"NB_API\s+(?:const\s+char\*|int)\s+(nb_\w+)\s*\(", text)))


def load():
    """Loads the product library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NBodyError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()) first")
    L = C.CDLL(LIB_PATH)
    vp, sz, f32, f64 = C.c_void_p, C.c_size_t, C.c_float, C.c_double
    L.nb_last_error.restype = C.c_char_p
    L.nb_default_config.argtypes = [C.POINTER(Config)]
    L.nb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.nb_destroy.argtypes = [vp]
    L.nb_set_theta.argtypes = [vp, f32]
    L.nb_init_aos.argtypes = [vp, vp, sz, sz]
    L.nb_init_soa.argtypes = [vp, vp, vp, vp, sz]
    L.nb_seed_galaxy_host.argtypes = [vp, sz, sz, C.c_uint64, f32]
    L.nb_seed_default_options.argtypes = [C.POINTER(SeedOptions)]
    L.nb_seed_host.argtypes = [C.c_int, vp, sz, sz, C.c_uint64, C.POINTER(SeedOptions)]
    L.nb_seed_collision_host.argtypes = [vp, sz, sz, C.c_uint64, f32, f32, f64]
    L.nb_seed_galaxy_device.argtypes = [vp, sz, C.c_uint64, f32]
    L.nb_seed_collision_device.argtypes = [vp, sz, C.c_uint64, f32, f32, f64]
    L.nb_seed_device.argtypes = [C.c_int, C.c_int, vp, sz, sz, C.c_uint64, vp]
    L.nb_get_aos_records.argtypes = [vp, vp, sz, vp]
    L.nb_scale_masses.argtypes = [vp, f64]
    L.nb_enable_graphs.argtypes = [vp, C.c_int]
    L.nb_step.argtypes = [vp, f32, C.c_int]
    L.nb_update_aos.argtypes = [vp, vp, sz, sz, f32]
    L.nb_sync.argtypes = [vp]
    L.nb_read_aos.argtypes = [vp, vp, sz, sz]
    L.nb_read_soa.argtypes = [vp, vp, vp]
    L.nb_owned_range.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.nb_num_bodies.argtypes = [vp, C.POINTER(sz)]
    L.nb_compute_accel.argtypes = [vp]
    L.nb_get_accel.argtypes = [vp, vp]
    L.nb_get_accel_of.argtypes = [vp, vp, sz, vp]
    L.nb_get_step_accel_of.argtypes = [vp, vp, sz, vp]
    L.nb_direct_accel.argtypes = [vp, vp, sz, vp]
    L.nb_state_hash.argtypes = [vp, vp]
    L.nb_get_morton.argtypes = [vp, vp, vp, C.POINTER(sz)]
    L.nb_get_tree.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(sz)]
    L.nb_get_walk_stats.argtypes = [vp, vp]
    L.nb_get_walk_sparse_load.argtypes = [vp, vp]
    L.nb_get_walk_occupancy.argtypes = [vp, vp]
    L.nb_get_leaf_cells.argtypes = [vp, vp, vp, C.POINTER(sz)]
    L.nb_energy.argtypes = [vp, C.POINTER(f64), C.POINTER(f64)]
    L.nb_energy_sampled.argtypes = [vp, sz, C.POINTER(f64), C.POINTER(f64), C.POINTER(sz)]
    L.nb_closest_particle.argtypes = [vp, C.POINTER(f32), C.POINTER(sz), C.POINTER(f32)]
    L.nb_nbody_save.argtypes = [C.c_char_p, vp, sz, sz]
    L.nb_nbody_count.argtypes = [C.c_char_p, C.POINTER(sz)]
    L.nb_nbody_load.argtypes = [C.c_char_p, vp, sz, sz, C.POINTER(sz), C.c_int]
    L.nb_nbody_recentre.argtypes = [vp, sz, sz]
    L.nb_comm_unique_id.argtypes = [vp]
    L.nb_comm_init.argtypes = [vp, vp]
    L.nb_device_posw.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.nb_mark_exchanged.argtypes = [vp]
    L.nb_p2p_export.argtypes = [vp, vp]
    L.nb_p2p_attach.argtypes = [vp, vp]
    L.nb_p2p_attach_local.argtypes = [vp, C.POINTER(vp)]
    L.nb_last_step_timing.argtypes = [vp, C.POINTER(f32), C.POINTER(f32), C.POINTER(C.c_int)]
    L.nb_step_timing_mean.argtypes = [vp, C.c_int, C.POINTER(f32), C.POINTER(f32), C.POINTER(C.c_int)]
    L.nb_step_period_mean.argtypes = [vp, C.c_int, C.POINTER(f32), C.POINTER(C.c_int)]
    L.nb_probe_fp32_peak.argtypes = [vp, C.POINTER(f64)]
    L.nb_last_build_timing.argtypes = [vp, C.POINTER(f32)]
    L.nb_shard_range.argtypes = [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise NBodyError(f"nb status {rc}: {load().nb_last_error().decode(errors='replace')}")


def seed_galaxy_host(n, seed=42, scale=1.0):
    """GalaxySeeder<Particle>(particles, scale).Seed(seed) -- reference GalaxySeeder.cpp:43-80."""
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    _check(load().nb_seed_galaxy_host(p.ctypes.data, n, PARTICLE_DTYPE.itemsize, seed, scale))
    return p


def seed_host(kind, n, seed=0, scale=1.0, colours=None, lw=False):
    """CreateParticleSeeder<T>(particles, kind, scale)->Seed(seed) -- reference IParticleSeeder.hpp:29-50 --
    for T = Particle or (lw=True) T = LWParticle; colours = ((r_lo, r_hi), (g_lo, g_hi), (b_lo, b_hi))
    are the GalaxySeeder Set{Red,Green,Blue}Dist ranges."""
    dtype = LWPARTICLE_DTYPE if lw else PARTICLE_DTYPE
    p = np.zeros(n, dtype=dtype)
    o = SeedOptions()
    _check(load().nb_seed_default_options(C.byref(o)))
    o.layout = LAYOUT_LWPARTICLE if lw else LAYOUT_PARTICLE
    o.scale = scale
    if colours is not None:
        (o.red[0], o.red[1]), (o.green[0], o.green[1]), (o.blue[0], o.blue[1]) = colours
    _check(load().nb_seed_host(kind, p.ctypes.data, n, dtype.itemsize, seed, C.byref(o)))
    return p


def seed_device(kind, n, seed=0, scale=1.0, colours=None, lw=False, device=0):
    """The same seeders run ON THE DEVICE (csrc/seed_device.cu: parallel parse of the minstd_rand0 stream), bit for
    bit equal to seed_host / the reference; records are copied back to a host array."""
    dtype = LWPARTICLE_DTYPE if lw else PARTICLE_DTYPE
    p = np.zeros(n, dtype=dtype)
    o = SeedOptions()
    _check(load().nb_seed_default_options(C.byref(o)))
    o.layout = LAYOUT_LWPARTICLE if lw else LAYOUT_PARTICLE
    o.scale = scale
    if colours is not None:
        (o.red[0], o.red[1]), (o.green[0], o.green[1]), (o.blue[0], o.blue[1]) = colours
    _check(load().nb_seed_device(kind, device, p.ctypes.data, n, dtype.itemsize, seed, C.byref(o)))
    return p


def seed_collision_host(n, seed=42, scale=1.0, separation=2000.0, approach_speed=2e16):
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    _check(load().nb_seed_collision_host(p.ctypes.data, n, PARTICLE_DTYPE.itemsize, seed, scale,
                                         separation, approach_speed))
    return p


def save_nbody(path, particles):
    """Writes the reference's .nbody file (raw 104-byte Particle records, SimulationState.cpp:317-331)."""
    p = np.ascontiguousarray(particles)
    _check(load().nb_nbody_save(os.fsencode(path), p.ctypes.data, len(p), p.dtype.itemsize))


def load_nbody(path, recentre=True):
    """InitParticlesFromFile (SimulationState.cpp:229-277): all whole records, recentred on the centre of mass."""
    n = C.c_size_t()
    _check(load().nb_nbody_count(os.fsencode(path), C.byref(n)))
    p = np.zeros(n.value, dtype=PARTICLE_DTYPE)
    got = C.c_size_t()
    _check(load().nb_nbody_load(os.fsencode(path), p.ctypes.data, len(p), PARTICLE_DTYPE.itemsize, C.byref(got), int(recentre)))
    return p[: got.value]


def recentre(particles):
    _check(load().nb_nbody_recentre(particles.ctypes.data, len(particles), particles.dtype.itemsize))
    return particles


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap a device pointer the library owns."""

    def __init__(self, ptr, nbytes, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {
            "shape": (nbytes // 4,),
            "typestr": "<f4",
            "data": (ptr, False),
            "version": 2,
        }


class Sim:
    """One engine handle == one INBodySim instance of the reference (Init / Update / read back)."""

    def __init__(self, mode=MODE_ALLPAIRS, theta=2.0, device=0, rank=0, world=1, stream=None,
                 source_splits=0, kernel_variant=0):
        L = load()
        cfg = Config()
        _check(L.nb_default_config(C.byref(cfg)))
        cfg.mode, cfg.theta, cfg.device = mode, theta, device
        cfg.rank, cfg.world = rank, world
        cfg.stream = stream
        cfg.source_splits, cfg.kernel_variant = source_splits, kernel_variant
        self.cfg = cfg
        self._h = C.c_void_p()
        _check(L.nb_create(C.byref(cfg), C.byref(self._h)))
        self._L = L

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.nb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- INBodySim::Init
    def init(self, particles):
        assert particles.dtype.itemsize >= 104 and particles.flags["C_CONTIGUOUS"]
        _check(self._L.nb_init_aos(self._h, particles.ctypes.data, len(particles), particles.dtype.itemsize))
        self.n = len(particles)

    def init_soa(self, pos, vel, mass):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        _check(self._L.nb_init_soa(self._h, pos.ctypes.data, vel.ctypes.data, mass.ctypes.data, len(mass)))
        self.n = len(mass)

    def seed_galaxy_device(self, n, seed=42, scale=1.0):
        _check(self._L.nb_seed_galaxy_device(self._h, n, seed, scale))
        self.n = n

    def seed_collision_device(self, n, seed=42, scale=1.0, separation=2000.0, approach_speed=2e16):
        _check(self._L.nb_seed_collision_device(self._h, n, seed, scale, separation, approach_speed))
        self.n = n

    def enable_graphs(self, on):
        _check(self._L.nb_enable_graphs(self._h, 1 if on else 0))

    def scale_masses(self, factor):
        _check(self._L.nb_scale_masses(self._h, factor))

    def aos_records(self, bodies):
        """Records of the device image of the Particle array for the listed bodies."""
        b = np.ascontiguousarray(bodies, dtype=np.uint32)
        out = np.zeros(len(b), dtype=PARTICLE_DTYPE)
        _check(self._L.nb_get_aos_records(self._h, b.ctypes.data, len(b), out.ctypes.data))
        return out

    # --- INBodySim::Update
    def step(self, dt, nsteps=1):
        _check(self._L.nb_step(self._h, dt, nsteps))

    def update(self, particles, dt):
        """Update(dt) with the reference's host-array contract (upload, step, write back)."""
        _check(self._L.nb_update_aos(self._h, particles.ctypes.data, len(particles), particles.dtype.itemsize, dt))

    def sync(self):
        _check(self._L.nb_sync(self._h))

    def set_theta(self, theta):
        _check(self._L.nb_set_theta(self._h, theta))

    # --- read back
    def read(self, particles):
        _check(self._L.nb_read_aos(self._h, particles.ctypes.data, len(particles), particles.dtype.itemsize))
        return particles

    def owned_range(self):
        a, b = C.c_size_t(), C.c_size_t()
        _check(self._L.nb_owned_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def read_soa(self):
        _, cnt = self.owned_range()
        pos = np.zeros((cnt, 3), dtype=np.float32)
        vel = np.zeros((cnt, 3), dtype=np.float64)
        _check(self._L.nb_read_soa(self._h, pos.ctypes.data, vel.ctypes.data))
        return pos, vel

    def accelerations(self):
        _, cnt = self.owned_range()
        _check(self._L.nb_compute_accel(self._h))
        acc = np.zeros((cnt, 3), dtype=np.float64)
        _check(self._L.nb_get_accel(self._h, acc.ctypes.data))
        return acc

    def accelerations_of(self, bodies):
        """Accelerations (current positions) of the listed global body indices; NaN rows for bodies not owned."""
        b = np.ascontiguousarray(bodies, dtype=np.uint32)
        acc = np.zeros((len(b), 3), dtype=np.float64)
        _check(self._L.nb_get_accel_of(self._h, b.ctypes.data, len(b), acc.ctypes.data))
        return acc

    def step_accelerations_of(self, bodies):
        """Accelerations the LAST step applied (at its pre-drift positions) of the listed global body indices."""
        b = np.ascontiguousarray(bodies, dtype=np.uint32)
        acc = np.zeros((len(b), 3), dtype=np.float64)
        _check(self._L.nb_get_step_accel_of(self._h, b.ctypes.data, len(b), acc.ctypes.data))
        return acc

    def direct_accelerations(self, bodies):
        """The reference's all-pairs law restated on the device for the listed bodies x all sources."""
        b = np.ascontiguousarray(bodies, dtype=np.uint32)
        acc = np.zeros((len(b), 3), dtype=np.float64)
        _check(self._L.nb_direct_accel(self._h, b.ctypes.data, len(b), acc.ctypes.data))
        return acc

    def state_hash(self):
        """(checksum of all positions, checksum of the owned velocities)."""
        h = np.zeros(2, dtype=np.uint64)
        _check(self._L.nb_state_hash(self._h, h.ctypes.data))
        return int(h[0]), int(h[1])

    def morton(self):
        codes = np.zeros(self.n, dtype=np.uint64)
        order = np.zeros(self.n, dtype=np.uint32)
        m = C.c_size_t()
        _check(self._L.nb_get_morton(self._h, codes.ctypes.data, order.ctypes.data, C.byref(m)))
        return codes[: m.value], order[: m.value]

    def inbounds(self):
        """Bodies inside the root cube at the last tree build (the others are dropped as sources, Octree.cpp:58-62)."""
        m = C.c_size_t()
        _check(self._L.nb_get_morton(self._h, None, None, C.byref(m)))
        return m.value

    def tree(self):
        m = C.c_size_t()
        _check(self._L.nb_get_tree(self._h, None, None, None, None, None, C.byref(m)))
        k = m.value
        left = np.zeros(k, dtype=np.int32)
        right = np.zeros(k, dtype=np.int32)
        prefix = np.zeros(k, dtype=np.int32)
        mass = np.zeros(k, dtype=np.float64)
        com = np.zeros((k, 3), dtype=np.float32)
        _check(self._L.nb_get_tree(self._h, left.ctypes.data, right.ctypes.data, prefix.ctypes.data,
                                   mass.ctypes.data, com.ctypes.data, C.byref(m)))
        return dict(left=left, right=right, prefix=prefix, mass=mass, com=com)

    def walk_stats(self):
        s = np.zeros(3, dtype=np.uint64)
        _check(self._L.nb_get_walk_stats(self._h, s.ctypes.data))
        return dict(cell_evals=int(s[0]), leaf_evals=int(s[1]), visits=int(s[2]))

    def walk_occupancy(self):
        """Lane-occupancy histogram [0..32] of the traversal walk_stats() just ran."""
        h = (C.c_uint64 * 33)()
        _check(self._L.nb_get_walk_occupancy(self._h, h))
        return np.array(list(h), dtype=np.uint64)

    def walk_sparse_load(self):
        out = np.zeros(7, dtype=np.uint64)
        _check(self._L.nb_get_walk_sparse_load(self._h, out.ctypes.data))
        return out

    def energy(self):
        ke, pe = C.c_double(), C.c_double()
        _check(self._L.nb_energy(self._h, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    def leaf_cells(self):
        """(cubes[k] = {centre.xyz, size}, body[k]) of the occupied octree leaves, Morton order."""
        n = C.c_size_t()
        _check(self._L.nb_get_leaf_cells(self._h, None, None, C.byref(n)))
        cells = np.zeros((n.value, 4), dtype=np.float32)
        body = np.zeros(n.value, dtype=np.uint32)
        _check(self._L.nb_get_leaf_cells(self._h, cells.ctypes.data, body.ctypes.data, C.byref(n)))
        return cells, body

    def closest_particle(self, pos):
        """Maths::ClosestParticle on the device-resident positions -> (index, distance squared)."""
        q = (C.c_float * 3)(*[float(x) for x in pos])
        idx, d = C.c_size_t(), C.c_float()
        _check(self._L.nb_closest_particle(self._h, q, C.byref(idx), C.byref(d)))
        return idx.value, d.value

    def energy_sampled(self, stride):
        """(kinetic, potential estimate, samples): see nb_energy_sampled."""
        ke, pe, ns = C.c_double(), C.c_double(), C.c_size_t()
        _check(self._L.nb_energy_sampled(self._h, C.c_size_t(stride), C.byref(ke), C.byref(pe), C.byref(ns)))
        return ke.value, pe.value, ns.value

    # --- multi-GPU plumbing
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        _check(load().nb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self._L.nb_comm_init(self._h, buf))

    def device_posw(self):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _check(self._L.nb_device_posw(self._h, C.byref(ptr), C.byref(nbytes)))
        return _CudaArray(ptr.value, nbytes.value, self)

    P2P_HANDLE_BYTES = 384

    def p2p_export(self):
        buf = (C.c_uint8 * self.P2P_HANDLE_BYTES)()
        _check(self._L.nb_p2p_export(self._h, buf))
        return bytes(buf)

    def p2p_attach(self, all_handles):
        blob = b"".join(all_handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(self._L.nb_p2p_attach(self._h, buf))

    def p2p_attach_local(self, sims):
        arr = (C.c_void_p * len(sims))(*[s._h for s in sims])
        _check(self._L.nb_p2p_attach_local(self._h, arr))

    def mark_exchanged(self):
        _check(self._L.nb_mark_exchanged(self._h))

    # --- measurement
    def last_step_timing(self):
        t, f, k = C.c_float(), C.c_float(), C.c_int()
        _check(self._L.nb_last_step_timing(self._h, C.byref(t), C.byref(f), C.byref(k)))
        return t.value, f.value, k.value

    def step_timing_mean(self, max_steps=0):
        """(dominant kernel ms, tree build ms, steps averaged) over the last force passes (<= 64 kept)."""
        a, b, k = C.c_float(), C.c_float(), C.c_int()
        _check(self._L.nb_step_timing_mean(self._h, max_steps, C.byref(a), C.byref(b), C.byref(k)))
        return a.value, b.value, k.value

    def step_period_mean(self, max_steps=0):
        """(mean device period of the last steps in ms, steps averaged)."""
        a, k = C.c_float(), C.c_int()
        _check(self._L.nb_step_period_mean(self._h, max_steps, C.byref(a), C.byref(k)))
        return a.value, k.value

    def last_build_ms(self):
        t = C.c_float()
        _check(self._L.nb_last_build_timing(self._h, C.byref(t)))
        return t.value

    def probe_fp32_peak(self):
        v = C.c_double()
        _check(self._L.nb_probe_fp32_peak(self._h, C.byref(v)))
        return v.value
