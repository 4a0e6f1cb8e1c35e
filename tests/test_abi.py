"""C-ABI surface: the library loads and exports every symbol include/nbody_b200.h declares; the
host-only entry points work; without a GPU the compute entry points fail loudly.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from conftest import as_particles, load_golden, same_particles


def test_exports_every_declared_symbol(pkg):
    lib = pkg.load()
    names = pkg.declared_symbols()
    assert len(names) >= 30
    missing = [s for s in names if not hasattr(lib, s)]
    assert missing == []
    assert lib.nb_abi_version() == 1


def test_default_config_holds_reference_constants(pkg):
    cfg = pkg.Config()
    assert pkg.load().nb_default_config(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(pkg.Config)
    assert cfg.G == 6.674e-11 and cfg.softening == 10.0            # Physics.hpp:9-10
    assert cfg.position_scale == 20 * 1.15e12                      # Physics.hpp:13,16
    assert cfg.bounds == 4000.0 and cfg.theta == 2.0               # BarnesHut.cpp:14, Octree.cpp:5
    assert cfg.world == 1 and cfg.rank == 0


def test_bad_arguments_are_rejected(pkg):
    lib = pkg.load()
    assert lib.nb_default_config(None) == -1
    cfg = pkg.Config()
    lib.nb_default_config(C.byref(cfg))
    cfg.struct_size = 8
    h = C.c_void_p()
    assert lib.nb_create(C.byref(cfg), C.byref(h)) == -1
    assert b"size mismatch" in lib.nb_last_error()
    assert lib.nb_step(None, 0.01, 1) == -1
    buf = np.zeros(10, dtype=pkg.PARTICLE_DTYPE)
    assert lib.nb_seed_galaxy_host(buf.ctypes.data, 10, 100, 1, 1.0) == -1      # stride < 104
    # the device seeders validate before they touch the GPU
    assert lib.nb_seed_device(pkg.SEEDER_GALAXY, 0, buf.ctypes.data, 10, 100, 1, None) == -1          # stride < 104
    assert lib.nb_seed_device(7, 0, buf.ctypes.data, 10, 104, 1, None) == -1 and b"unknown seeder" in lib.nb_last_error()
    assert lib.nb_seed_device(pkg.SEEDER_STARSYSTEM, 0, buf.ctypes.data, 0, 104, 1, None) == -1
    assert lib.nb_seed_device(pkg.SEEDER_GALAXY, 0, None, 10, 104, 1, None) == -1
    assert lib.nb_seed_device(pkg.SEEDER_GALAXY, 0, None, 0, 104, 1, None) == 0                       # nothing to seed
    for fn in (lib.nb_get_accel_of, lib.nb_get_step_accel_of, lib.nb_direct_accel):
        assert fn(None, None, 0, None) == -1
    assert lib.nb_state_hash(None, None) == -1 and lib.nb_scale_masses(None, 2.0) == -1 and lib.nb_enable_graphs(None, 1) == -1


def test_no_gpu_the_device_seeder_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.NBodyError):
        pkg.seed_device(pkg.SEEDER_GALAXY, 100, 1)


def test_no_gpu_means_error_not_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.NBodyError):
        pkg.Sim()


def test_host_seeder_matches_reference_golden(pkg):
    seeds = load_golden("galaxy_seeds.npz")
    for key, (n, seed, scale) in {"n256_s42": (256, 42, 1.0), "n1000_s7": (1000, 7, 2.5), "n4096_s42": (4096, 42, 1.0)}.items():
        want = as_particles(seeds[key], pkg.PARTICLE_DTYPE)
        got = pkg.seed_galaxy_host(n, seed, scale)
        assert same_particles(got, want), key


def test_host_seeder_layout(pkg):
    """Appendix A.5: bodies [0, 2*61*floor(0.4 N / 60)) are arm bodies of mass 1e20."""
    n = 4096
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    arms = 2 * 61 * int(np.floor(np.float32(n) * np.float32(0.4) / 60))
    assert arms == 3294
    assert np.all(p["Mass"][:arms] == 1e20)
    assert np.all((p["Mass"][arms:] >= 1e28) & (p["Mass"][arms:] <= 1e30))
    assert np.all(p["Forces"] == 0)
    assert np.array_equal(p["Colour"], p["OriginalColour"])
    assert np.all(p["Velocity"][:, 2] == 0)     # velocities are not rotated (GalaxySeeder.cpp:91-93)


def test_collision_scene(pkg):
    n = 2000
    p = pkg.seed_collision_host(n, 42, 1.0, separation=2000.0, approach_speed=2e16)
    a = pkg.seed_galaxy_host(n // 2, 42, 1.0)
    b = pkg.seed_galaxy_host(n // 2, 43, 1.0)
    assert np.array_equal(p["Mass"], np.concatenate([a["Mass"], b["Mass"]]))
    assert np.allclose(p["Position"][: n // 2, 0], a["Position"][:, 0] - 1000.0)
    assert np.allclose(p["Position"][n // 2:, 0], b["Position"][:, 0] + 1000.0)
    assert np.allclose(p["Velocity"][: n // 2, 0], a["Velocity"][:, 0] + 2e16)
    assert np.all(np.abs(p["Position"]) < 4000.0)
