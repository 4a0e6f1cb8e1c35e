"""The C++ adapter `B200Sim : INBodySim` (procedural-universe_b200/host) driven through the
reference's own interface and factory, next to the reference's sims driven the same way
(oracle/adapter_driver.cpp, built by oracle/build_ref.sh where the reference tree exists)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, as_particles, load_golden, rel_err, same_particles

LIB = os.path.join(ROOT, "oracle", "_ref", "libb200_adapter_test.so")
needs_lib = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libb200_adapter_test.so not built")

REFERENCE, B200 = 0, 1
BRUTE_CPU, BRUTE_GPU, BARNES_HUT = 0, 1, 2     # ENBodySim, INBodySim.hpp:11-17


def run(p, impl, sim_type, dt, steps, theta=0.5, recolour=False):
    lib = C.CDLL(LIB)
    lib.adapter_run.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int,
                                C.c_char_p, C.c_size_t]
    q = p.copy()
    log = C.create_string_buffer(512)
    rc = lib.adapter_run(q.ctypes.data, len(q), impl, sim_type, dt, steps, theta, int(recolour), log, 512)
    return rc, q, log.value.decode()


@needs_lib
def test_reference_sims_through_the_interface_match_golden(pkg):
    g = load_golden("allpairs_n256.npz")
    p = as_particles(load_golden("galaxy_seeds.npz")["n256_s42"], pkg.PARTICLE_DTYPE).copy()
    rc, q, log = run(p, REFERENCE, BRUTE_CPU, float(g["dt"]), int(g["steps"]))
    assert rc == 0 and log == "[Info] Brute Force CPU\n"          # BruteForceCPU.cpp:17, format of test/LogTests.cpp
    assert same_particles(q, as_particles(g["state10"], pkg.PARTICLE_DTYPE))
    g = load_golden("barneshut_n1024.npz")
    p = pkg.seed_galaxy_host(1024, 42, 1.0)
    rc, q, log = run(p, REFERENCE, BARNES_HUT, float(g["dt"]), 5, theta=0.5)
    assert rc == 0 and log == "[Info] Barnes-Hut\n"
    assert same_particles(q, as_particles(g["state5"], pkg.PARTICLE_DTYPE))


@needs_lib
def test_adapter_without_gpu_logs_and_leaves_particles_alone(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = pkg.seed_galaxy_host(64, 1, 1.0)
    rc, q, log = run(p, B200, BRUTE_GPU, 0.01, 2)
    assert rc == 0
    assert log.startswith("[Info] Brute Force B200\n[Error] B200 engine unavailable")
    assert same_particles(q, p)                                   # no CPU fallback: nothing moved


@needs_lib
@pytest.mark.gpu
def test_adapter_allpairs_equals_reference_bruteforce(pkg):
    p = pkg.seed_galaxy_host(2048, 8, 1.0)
    rc0, want, _ = run(p, REFERENCE, BRUTE_CPU, 0.01, 5)
    rc1, got, log = run(p, B200, BRUTE_GPU, 0.01, 5, recolour=True)
    assert rc0 == 0 and rc1 == 0 and log == "[Info] Brute Force B200\n"
    dv = want["Velocity"] - p["Velocity"]
    err = np.linalg.norm(got["Velocity"] - want["Velocity"], axis=1) / np.linalg.norm(dv, axis=1)
    assert np.median(err) < 1e-5 and err.max() < 1e-3
    assert np.abs(got["Position"] - want["Position"]).max() < 1e-2
    assert np.all(got["Forces"] == 0)                             # BruteForceCPU.cpp:72
    # the caller recoloured particles 0..4 between Updates; the engine never writes colours
    assert np.allclose(got["Colour"][:5, :3], [0.5, 0.25, 0.125])
    assert np.array_equal(got["Colour"][5:], p["Colour"][5:])
    assert np.array_equal(got["OriginalColour"], p["OriginalColour"]) and np.array_equal(got["Mass"], p["Mass"])


@needs_lib
@pytest.mark.gpu
def test_adapter_barneshut_equals_reference_barneshut(pkg):
    p = pkg.seed_galaxy_host(4096, 8, 1.0)
    dt = np.float32(0.02 / 60)
    rc0, want, _ = run(p, REFERENCE, BARNES_HUT, dt, 3, theta=0.5)
    rc1, got, log = run(p, B200, BARNES_HUT, dt, 3, theta=0.5)
    assert rc0 == 0 and rc1 == 0 and log == "[Info] Barnes-Hut B200\n"
    dv = want["Velocity"] - p["Velocity"]
    err = np.linalg.norm(got["Velocity"] - want["Velocity"], axis=1) / np.linalg.norm(dv, axis=1)
    assert np.median(err) < 1e-3
    assert np.median(rel_err(got["Forces"], want["Forces"])) < 1e-3     # BarnesHut leaves the last force in place
    assert np.abs(got["Position"] - want["Position"]).max() < 1e-3


def debug_cubes(p, impl, dt, steps, theta=0.5):
    lib = C.CDLL(LIB)
    lib.adapter_debug_cubes.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_size_t]
    lib.adapter_debug_cubes.restype = C.c_long
    q = p.copy()
    out = np.zeros((len(p), 4), dtype=np.float32)
    k = lib.adapter_debug_cubes(q.ctypes.data, len(q), impl, dt, steps, theta, out.ctypes.data, len(out))
    return k, out[:max(k, 0)]


@needs_lib
def test_reference_render_debug_draws_one_cube_per_occupied_leaf(pkg):
    from oracle import ref
    p = pkg.seed_galaxy_host(1024, 42, 1.0)
    k, cubes = debug_cubes(p, REFERENCE, np.float32(0.02 / 60), 1)
    want, _ = ref.octree_leaf_cubes(p)                           # the tree Update built: positions BEFORE the drift
    assert k == len(want) and np.array_equal(cubes, want)


@needs_lib
@pytest.mark.gpu
def test_adapter_render_debug_draws_the_reference_cubes(pkg):
    """Sim->RenderDebug(view, proj) through INBodySim on both sims: same cubes in the same order."""
    p = pkg.seed_galaxy_host(4096, 8, 1.0)
    dt = np.float32(0.02 / 60)
    k0, want = debug_cubes(p, REFERENCE, dt, 2)
    k1, got = debug_cubes(p, B200, dt, 2)
    assert k0 == k1 > 0
    assert np.array_equal(got[:, 3], want[:, 3])
    assert np.abs(got[:, :3] - want[:, :3]).max() <= 1e-3
