"""The reference's particle seeders ON THE DEVICE (csrc/seed_device.cu) against the reference itself.

GalaxySeeder<T>::Seed (reference src/Sim/GalaxySeeder.cpp:43-143) is one serial minstd_rand0 stream with
data-dependent draw counts; the device path parses that stream in parallel.  Every comparison here is byte for
byte: against `oracle/_ref` (the reference's own seeders compiled headless) where it was prebuilt, else against
the product's host seeder, which tests/test_seeders.py pins to the reference on CPU.
"""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, same_particles

pytestmark = pytest.mark.gpu


def _expect(pkg, kind, n, seed, scale=1.0, colours=None, lw=False):
    from oracle import ref
    if ref.available():
        return ref.seed_ex(n, kind, seed, scale, colours, lw)
    return pkg.seed_host(kind, n, seed, scale, colours, lw)


def _same(a, b):
    return all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


# 149: no arms (floor(0.4 n / 60) == 0); 150: one body per segment, ODD arm (61 bodies: the cached-variate
# alternation of distz flips between the arms); 4096: odd arm again; 1 and 2: disk only, ends inside a unit
@pytest.mark.parametrize("n", [1, 2, 3, 61, 149, 150, 151, 300, 4096, 5000, 65536, 100003])
@pytest.mark.parametrize("seed", [42, 43])
def test_galaxy_seeder_on_device_is_the_reference_stream(pkg, n, seed):
    got = pkg.seed_device(pkg.SEEDER_GALAXY, n, seed)
    want = _expect(pkg, pkg.SEEDER_GALAXY, n, seed)
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f]), (f, int(np.argmax((got[f] != want[f]).reshape(n, -1).any(axis=1))))


@pytest.mark.parametrize("seed", [0, 1, 7, 2147483647, 2**40 + 5])
def test_galaxy_seeder_seeds(pkg, seed):
    n = 20000
    assert _same(pkg.seed_device(pkg.SEEDER_GALAXY, n, seed), _expect(pkg, pkg.SEEDER_GALAXY, n, seed))


def test_galaxy_seeder_one_million_bodies(pkg):
    n = 1 << 20
    got = pkg.seed_device(pkg.SEEDER_GALAXY, n, 42)
    want = _expect(pkg, pkg.SEEDER_GALAXY, n, 42)
    assert _same(got, want)
    assert got.tobytes() == want.tobytes()          # padding bytes are zero on both sides


def test_galaxy_seeder_scale_colours_and_lwparticle(pkg):
    colours = ((0.1, 0.6), (0.0, 0.3), (0.5, 1.5))       # the blue range is clamped to [0.5, 1]
    for lw in (False, True):
        got = pkg.seed_device(pkg.SEEDER_GALAXY, 30000, 5, scale=2.5, colours=colours, lw=lw)
        assert _same(got, _expect(pkg, pkg.SEEDER_GALAXY, 30000, 5, 2.5, colours, lw))


@pytest.mark.parametrize("kind", ["SEEDER_RANDOM", "SEEDER_STARSYSTEM"])
def test_fixed_draw_seeders_on_device(pkg, kind):
    k = getattr(pkg, kind)
    for lw in (False, True):
        got = pkg.seed_device(k, 50000, 9, scale=3.0, lw=lw)
        assert _same(got, _expect(pkg, k, 50000, 9, 3.0, None, lw))


def test_seeding_the_handle_at_16m_matches_the_reference_golden(pkg):
    """nb_seed_galaxy_device at BASELINE's 2^24: the device image equals what the reference seeded
    (sha256 of the whole 1.7 GB array, recorded by tests/golden/make_golden_bh16m.py)."""
    g = load_golden("bh_bh16m_sampled.npz")
    n = int(g["n"])
    sim = pkg.Sim(mode=pkg.MODE_BARNESHUT, theta=0.5)
    sim.seed_galaxy_device(n, 42, 1.0)
    rec = sim.aos_records(g["targets"])
    want = np.ascontiguousarray(g["records"]).view(pkg.PARTICLE_DTYPE).reshape(-1)
    assert same_particles(rec, want)
    p = np.zeros(n, dtype=pkg.PARTICLE_DTYPE)
    sim.read(p)                                        # Position / Velocity / Forces of every body
    host = pkg.seed_galaxy_host(n, 42, 1.0)
    assert np.array_equal(p["Position"], host["Position"]) and np.array_equal(p["Velocity"], host["Velocity"])
    assert hashlib.sha256(host.view(np.uint8)).hexdigest() == str(g["sha256"])
    sim.close()


def test_collision_scene_on_device(pkg):
    n = 200001
    sim = pkg.Sim(mode=pkg.MODE_BARNESHUT, theta=0.5)
    sim.seed_collision_device(n, 42, 1.0, 2000.0, 2e16)
    host = pkg.seed_collision_host(n, 42, 1.0, 2000.0, 2e16)
    rec = sim.aos_records(np.arange(0, n, 97))
    assert same_particles(rec, host[::97])
    pos, vel = sim.read_soa()
    assert np.array_equal(pos, host["Position"]) and np.array_equal(vel, host["Velocity"])
    sim.close()
