"""The host seeders (nb_seed_host: RandomSeeder, GalaxySeeder, StarSystemSeeder for Particle and
LWParticle records) against golden vectors produced by the reference's own templates, and --
where the reference build is present -- against that build directly.  CPU only."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpu_ref.so not built")

_spec = importlib.util.spec_from_file_location("make_golden_seeders", os.path.join(GOLDEN, "make_golden_seeders.py"))
_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_gen)


def _same(a, b):
    return all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


def test_seeders_match_reference_golden(pkg):
    g = load_golden("seeders.npz")
    for kind, n, seed, scale, col, lw in _gen.CASES:
        dtype = pkg.LWPARTICLE_DTYPE if lw else pkg.PARTICLE_DTYPE
        want = np.ascontiguousarray(g[_gen.key(kind, n, seed, scale, col, lw)]).view(dtype).reshape(-1)
        got = pkg.seed_host(kind, n, seed=seed, scale=scale, colours=_gen.CASES_COLOURS if col else None, lw=lw)
        assert _same(got, want), (kind, n, seed, scale, col, lw)


@needs_ref
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("lw", [False, True])
def test_seeders_match_reference_build(pkg, kind, lw):
    for n in (1, 7, 1000, 20000):
        for scale in (1.0, 0.1, 4.0):
            for col in (None, ((0.1, 0.9), (0.0, 0.3), (0.5, 2.0))):
                got = pkg.seed_host(kind, n, seed=(2 << 21) + n, scale=scale, colours=col, lw=lw)
                want = ref.seed_ex(n, kind, (2 << 21) + n, scale, col, lw)
                assert _same(got, want), (kind, lw, n, scale, col)


def test_random_and_starsystem_ignore_their_seed(pkg):
    """RandomSeeder.cpp:15 and StarSystemSeeder.cpp:30 construct a fresh default engine."""
    for kind in (pkg.SEEDER_RANDOM, pkg.SEEDER_STARSYSTEM):
        assert _same(pkg.seed_host(kind, 50, seed=1), pkg.seed_host(kind, 50, seed=99))
    assert not _same(pkg.seed_host(pkg.SEEDER_GALAXY, 50, seed=1), pkg.seed_host(pkg.SEEDER_GALAXY, 50, seed=99))


def test_starsystem_layout(pkg):
    """StarSystemSeeder.cpp:20-28, 41-46: a 1e30 star at the origin, the rest on the +z axis."""
    p = pkg.seed_host(pkg.SEEDER_STARSYSTEM, 200)
    assert p["Mass"][0] == 1e30 and np.all(p["Position"][0] == 0) and np.all(p["Velocity"][0] == 0)
    assert np.allclose(p["Colour"][0], [0.6, 1.0, 1.0, 1.0])
    z = p["Position"][1:, 2]
    assert np.all(p["Position"][1:, :2] == 0) and np.all((z >= 200.0 - 1e-3) & (z <= 350.0 + 1e-3))   # 4..7 AU*M / 2.3e13
    assert np.all(p["Velocity"][1:, 2] == 0)
    assert np.all((p["Mass"][1:] >= 1e10) & (p["Mass"][1:] <= 1e26))


def test_lw_records_carry_only_position_colour_scale(pkg):
    lw = pkg.seed_host(pkg.SEEDER_GALAXY, 500, seed=11, scale=0.1, lw=True)
    full = pkg.seed_host(pkg.SEEDER_GALAXY, 500, seed=11, scale=0.1, lw=False)
    assert np.array_equal(lw["Position"], full["Position"]) and np.array_equal(lw["Colour"], full["Colour"])
    assert np.all(lw["Scale"] == 1.0)


def test_seed_host_argument_checks(pkg):
    lib = pkg.load()
    buf = np.zeros(4, dtype=pkg.PARTICLE_DTYPE)
    assert lib.nb_seed_host(7, buf.ctypes.data, 4, 104, 0, None) == -1               # unknown seeder
    assert lib.nb_seed_host(1, buf.ctypes.data, 4, 100, 0, None) == -1               # stride below the record
    assert lib.nb_seed_host(2, buf.ctypes.data, 0, 104, 0, None) == -1               # star system needs a star
    o = pkg.SeedOptions()
    lib.nb_seed_default_options(C.byref(o))
    o.struct_size = 4
    assert lib.nb_seed_host(1, buf.ctypes.data, 4, 104, 0, C.byref(o)) == -1
    assert lib.nb_seed_host(1, None, 0, 104, 0, None) == 0                           # empty is fine
