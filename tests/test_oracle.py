"""The oracle restatement (oracle/nbody_port.c) against the golden vectors generated from the
reference's own code, and -- where the reference build is present -- against that build directly.
CPU only."""
import numpy as np
import pytest

from conftest import as_particles, load_golden, same_particles
from oracle import port, ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpu_ref.so not built")


def test_port_allpairs_matches_golden(particle_dtype):
    g = load_golden("allpairs_n256.npz")
    seeds = load_golden("galaxy_seeds.npz")
    p = as_particles(seeds["n256_s42"], particle_dtype)
    assert np.array_equal(port.allpairs_forces(p), g["forces0"])           # bit exact
    q = port.allpairs_run(p, float(g["dt"]), int(g["steps"]))
    assert same_particles(q, as_particles(g["state10"], particle_dtype))


def test_port_config1_matches_golden(particle_dtype):
    """BASELINE.json configs[0]: N=4096, dt=0.01 -- forces of a target sample, bit exact."""
    g = load_golden("allpairs_n4096_100steps.npz")
    p = as_particles(load_golden("galaxy_seeds.npz")["n4096_s42"], particle_dtype)
    f = np.concatenate([port.allpairs_forces(p, int(t), 1) for t in g["targets"]])
    assert np.array_equal(f, g["forces0"])


def test_port_barneshut_matches_golden(particle_dtype):
    g = load_golden("barneshut_n1024.npz")
    p = np.zeros(1024, dtype=particle_dtype)
    from importlib import import_module
    p = import_module("procedural-universe_b200").seed_galaxy_host(1024, 42, 1.0)
    f, work = port.barneshut_forces(p, np.arange(1024), 0.5, want_counters=True)
    assert np.array_equal(f, g["forces"])
    assert [work["cell_evals"], work["leaf_evals"], work["visits"]] == list(g["work"])
    depth, path, stats = port.octree_paths(p)
    assert np.array_equal(depth, g["leaf_depth"]) and np.array_equal(path, g["path"])
    assert stats["max_depth"] == int(g["max_depth"])
    q = port.barneshut_run(p, float(g["dt"]), 5, 0.5)
    assert same_particles(q, as_particles(g["state5"], particle_dtype))


def test_morton_codes_are_the_reference_octree_paths(particle_dtype):
    """The 63-bit code of a body, truncated to its leaf depth, is the (z,y,x) child-index path the
    reference's Octree::Add takes (Octree.cpp:25-47, 53-84)."""
    g = load_golden("barneshut_n1024.npz")
    from importlib import import_module
    p = import_module("procedural-universe_b200").seed_galaxy_host(1024, 42, 1.0)
    codes = port.morton(p)
    depth = g["leaf_depth"]
    inb = codes != port.MORTON_OUTSIDE
    assert np.array_equal(~inb, depth == -1)
    top = codes[inb] >> (np.uint64(3) * (21 - depth[inb]).astype(np.uint64))
    assert np.array_equal(top, g["path"][inb])


def test_morton_edges():
    B = 4000.0
    assert port.lib().port_morton_one(-B, -B, -B) == 0
    assert port.lib().port_morton_one(B, 0.0, 0.0) == int(port.MORTON_OUTSIDE)          # half open
    assert port.lib().port_morton_one(0.0, 0.0, np.nextafter(np.float32(-B), np.float32(-1e9))) == int(port.MORTON_OUTSIDE)
    top = np.nextafter(np.float32(B), np.float32(0))
    assert port.lib().port_morton_one(top, top, top) == (1 << 63) - 1
    # first digit: x -> bit 0, y -> bit 1, z -> bit 2
    assert port.lib().port_morton_one(1.0, -1.0, -1.0) >> 60 == 1
    assert port.lib().port_morton_one(-1.0, 1.0, -1.0) >> 60 == 2
    assert port.lib().port_morton_one(-1.0, -1.0, 1.0) >> 60 == 4
    assert port.lib().port_morton_one(float("nan"), 0.0, 0.0) == int(port.MORTON_OUTSIDE)


def test_karras_tree_is_a_valid_radix_tree(particle_dtype):
    from importlib import import_module
    p = import_module("procedural-universe_b200").seed_galaxy_host(1024, 42, 1.0)
    codes, order = port.morton_sorted(p)
    m = len(codes)
    assert np.all(codes[1:] >= codes[:-1])
    left, right, prefix = port.karras(codes)
    seen_internal = np.zeros(m - 1, dtype=int)
    seen_leaf = np.zeros(m, dtype=int)
    for arr in (left, right):
        for c in arr:
            if c >= 0:
                seen_internal[c] += 1
            else:
                seen_leaf[~c] += 1
    assert seen_internal[0] == 0 and np.all(seen_internal[1:] == 1) and np.all(seen_leaf == 1)

    def span(node):
        if node < 0:
            return ~node, ~node
        a, _ = span(left[node])
        _, b = span(right[node])
        return a, b

    import sys
    sys.setrecursionlimit(10000)
    for i in range(0, m - 1, 37):
        a, b = span(i)
        x = int(codes[a]) ^ int(codes[b])
        assert prefix[i] == (64 - x.bit_length() if x else prefix[i])


@needs_ref
def test_port_equals_reference_build():
    p = ref.seed(1500, ref.SEED_GALAXY, 99, 1.0)
    assert np.array_equal(ref.bruteforce_forces(p, np.arange(0, 1500, 3)),
                          np.concatenate([port.allpairs_forces(p, t, 1) for t in range(0, 1500, 3)]))
    t = np.arange(0, 1500, 5)
    assert np.array_equal(ref.barneshut_forces(p, t, 0.5)[0], port.barneshut_forces(p, t, 0.5))
    assert np.array_equal(ref.barneshut_forces(p, t, 2.0)[0], port.barneshut_forces(p, t, 2.0))
    a = ref.barneshut_run(p, np.float32(0.02 / 60), 3, 0.5)[0]
    b = port.barneshut_run(p, np.float32(0.02 / 60), 3, 0.5)
    assert same_particles(a, b)
    # 1500 is not a multiple of 4: the driver picks a worker count that divides n
    a, _, w = ref.bruteforce_run(p, 0.01, 3)
    assert 1500 % w == 0
    assert same_particles(a, port.allpairs_run(p, 0.01, 3))


@needs_ref
def test_reference_theta_event_plumbing():
    """BHThetaChanged -> Octree::Theta (BarnesHut.cpp:29-31)."""
    ref.lib().ref_report_theta(0.75)
    assert ref.lib().ref_get_theta() == 0.75
