"""Barnes-Hut CUDA path against the oracle, through the C ABI.  GPU only (pytest -m gpu).

BASELINE.json north_star: Morton codes and tree topology bit-exact (against the CPU restatement,
which tests/test_oracle.py ties to the reference octree); accelerations within 1e-3 MEDIAN relative
at matched theta."""
import numpy as np
import pytest

from conftest import as_particles, load_golden, rel_err
from oracle import checker, port, ref

pytestmark = pytest.mark.gpu

BH_MEDIAN_RTOL = 1e-3


def bh(pkg, theta=0.5, **kw):
    return pkg.Sim(mode=pkg.MODE_BARNESHUT, theta=theta, **kw)


@pytest.fixture(scope="module")
def galaxy(pkg):
    return pkg.seed_galaxy_host(1024, 42, 1.0)


def scattered(pkg, n, seed, spread=3000.0, outside=0.1, duplicates=True):
    """Bodies filling the root cube, a fraction outside it, some on cell boundaries, some duplicated."""
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=pkg.PARTICLE_DTYPE)
    p["Position"] = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    k = int(outside * n)
    p["Position"][:k] *= 2.0                                   # many of these leave [-4000, 4000)
    p["Position"][k:k + 8] = [[0, 0, 0], [-4000, -4000, -4000], [2000, -2000, 1000], [4000, 0, 0],
                              [1.5, 1.5, 1.5], [1.5, 1.5, 1.5], [-0.0, 125.0, 3999.9998], [1e-30, -1e-30, 0]]
    if not duplicates:
        p["Position"][k + 5] = [1.5, 1.5, 1.75]
    p["Mass"] = rng.uniform(1e28, 1e30, n)
    p["Velocity"] = rng.normal(0, 1e15, (n, 3))
    return p


def test_morton_codes_and_order_bit_exact(pkg, galaxy):
    for p in (galaxy, scattered(pkg, 5000, 1), scattered(pkg, 4097, 2, spread=10.0)):
        sim = bh(pkg)
        sim.init(p)
        codes, order = sim.morton()
        want_codes, want_order = port.morton_sorted(p)
        assert len(codes) == len(want_codes)
        assert np.array_equal(codes, want_codes)
        assert np.array_equal(order, want_order)               # stable: ties keep body order
        sim.close()


def test_morton_codes_match_reference_octree_paths(pkg, galaxy):
    g = load_golden("barneshut_n1024.npz")
    sim = bh(pkg)
    sim.init(galaxy)
    codes, order = sim.morton()
    depth = g["leaf_depth"][order]
    top = codes >> (np.uint64(3) * (21 - depth).astype(np.uint64))
    assert np.array_equal(top, g["path"][order])
    sim.close()


def test_radix_tree_topology_bit_exact(pkg, galaxy):
    for p in (galaxy, scattered(pkg, 5000, 3), scattered(pkg, 777, 4, spread=1.0)):
        sim = bh(pkg)
        sim.init(p)
        codes, _ = sim.morton()
        t = sim.tree()
        left, right, prefix = port.karras(codes)
        assert np.array_equal(t["left"], left)
        assert np.array_equal(t["right"], right)
        assert np.array_equal(t["prefix"], prefix)
        sim.close()


def test_node_masses_and_centres_match_reference_cells(pkg, galaxy):
    """Radix-tree nodes that own an octree cell carry that cell's mass / centre of mass
    (Octree::CalculateMass, Octree.cpp:86-105; the reference accumulates the centre in fp32)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("needs oracle/_ref")
    sim = bh(pkg)
    sim.init(galaxy)
    codes, order = sim.morton()
    t = sim.tree()
    level = np.where(t["prefix"] >= 64, 21, np.minimum(21, (t["prefix"] - 1) // 3))
    parent_level = np.full(len(level), -1)
    for side in ("left", "right"):
        kids = t[side]
        internal = kids >= 0
        parent_level[kids[internal]] = level[internal]
    checked = 0
    for i in range(0, len(level), 3):
        if level[i] <= parent_level[i]:
            continue                                            # owns no octree cell (never visited)
        # first leaf under node i gives the cell's path
        node = i
        while node >= 0:
            node = t["left"][node]
        slot = ~node
        L = int(level[i])
        path = int(codes[slot]) >> (3 * (21 - L))
        cell = ref.octree_cell(galaxy, L, path)
        assert cell is not None
        mass, com, count, width = cell
        assert count >= 2 and width == 8000.0 / (1 << L)
        assert abs(t["mass"][i] - mass) <= 1e-6 * mass          # G m rounded to fp32 per body
        assert np.abs(t["com"][i] - com).max() <= 5e-3
        checked += 1
    assert checked > 100
    sim.close()


def test_accelerations_match_reference_theta_half(pkg, galaxy):
    g = load_golden("barneshut_n1024.npz")
    sim = bh(pkg, theta=0.5)
    sim.init(galaxy)
    acc = sim.accelerations()
    want = g["forces"] / galaxy["Mass"][:, None]
    err = rel_err(acc, want)
    print("BH theta=0.5 n=1024 rel err: median %.2e p99 %.2e max %.2e" % (np.median(err), np.quantile(err, 0.99), err.max()))
    assert np.median(err) < BH_MEDIAN_RTOL
    assert np.quantile(err, 0.99) < 1e-2
    # the same per-body acceptance rule => the same pair-evaluation count, up to cells right at width / r = theta
    # (the reference accumulates a cell's centre in fp32, this engine in fp64): within 0.2 %
    stats = sim.walk_stats()
    assert abs(stats["leaf_evals"] - int(g["work"][1])) <= 2e-3 * int(g["work"][1])
    sim.close()


@pytest.mark.parametrize("theta", [0.3, 1.0, 2.0])
def test_accelerations_other_thetas(pkg, theta):
    p = pkg.seed_galaxy_host(3000, 9, 1.0)
    sim = bh(pkg, theta=theta)
    sim.init(p)
    acc = sim.accelerations()
    t = np.arange(0, 3000, 3)
    want = checker.barneshut_accel(p, theta, t)
    err = rel_err(acc[t], want)
    print("BH theta=%.1f rel err: median %.2e p99 %.2e" % (theta, np.median(err), np.quantile(err, 0.99)))
    assert np.median(err) < BH_MEDIAN_RTOL
    sim.close()


def test_set_theta_matches_fresh_handle(pkg, galaxy):
    a = bh(pkg, theta=2.0)
    a.init(galaxy)
    a.accelerations()
    a.set_theta(0.5)                     # BHThetaChanged -> Octree::Theta (BarnesHut.cpp:29-31)
    b = bh(pkg, theta=0.5)
    b.init(galaxy)
    assert np.array_equal(a.accelerations(), b.accelerations())
    a.close()
    b.close()


def test_out_of_bounds_bodies(pkg):
    """Bodies outside the root cube are dropped as sources but still receive forces
    (Octree.cpp:58-62, BarnesHut.cpp:103-110)."""
    # no exactly coincident bodies here: the reference splits them until the cell size underflows
    # to zero and then loses both (and their mass) -- SURVEY.md appendix A, "do not replicate" 3;
    # this engine keeps them as a bucket at the finest Morton level (test_edge_cases).
    p = scattered(pkg, 4000, 5, duplicates=False)
    sim = bh(pkg, theta=0.5)
    sim.init(p)
    acc = sim.accelerations()
    t = np.arange(0, 4000, 4)
    want = checker.barneshut_accel(p, 0.5, t)
    ok = np.linalg.norm(want, axis=1) > 0
    err = rel_err(acc[t][ok], want[ok])
    assert np.median(err) < BH_MEDIAN_RTOL
    outside = (np.abs(p["Position"]) >= 4000).any(axis=1) | (p["Position"] < -4000).any(axis=1)
    assert outside.sum() > 100 and np.all(np.isfinite(acc))
    assert np.all(np.linalg.norm(acc[outside], axis=1) > 0)
    codes, order = sim.morton()
    assert len(codes) == 4000 - int(((p["Position"] < -4000) | (p["Position"] >= 4000)).any(axis=1).sum())
    sim.close()


def test_five_steps_against_golden(pkg, galaxy):
    g = load_golden("barneshut_n1024.npz")
    want = as_particles(g["state5"], pkg.PARTICLE_DTYPE)
    sim = bh(pkg, theta=0.5)
    sim.init(galaxy)
    sim.step(float(g["dt"]), 5)
    q = galaxy.copy()
    sim.read(q)
    dv_want = want["Velocity"] - galaxy["Velocity"]
    err = np.linalg.norm(q["Velocity"] - want["Velocity"], axis=1) / np.linalg.norm(dv_want, axis=1)
    print("BH 5 steps dv rel err: median %.2e p99 %.2e" % (np.median(err), np.quantile(err, 0.99)))
    assert np.median(err) < BH_MEDIAN_RTOL
    assert np.abs(q["Position"] - want["Position"]).max() < 1e-3
    # BarnesHut leaves the last force in Particle::Forces (BarnesHut.cpp:75,108)
    ferr = rel_err(q["Forces"], want["Forces"])
    assert np.median(ferr) < BH_MEDIAN_RTOL
    assert np.array_equal(q["Colour"], galaxy["Colour"])
    sim.close()


def test_edge_cases(pkg):
    dt = pkg.PARTICLE_DTYPE
    sim = bh(pkg, theta=0.5)
    # a single body; two bodies; everything outside the root; exact duplicates
    p = np.zeros(1, dtype=dt); p["Mass"] = 1e30
    sim.init(p)
    assert np.all(sim.accelerations() == 0)
    p = np.zeros(2, dtype=dt); p["Mass"] = (1e30, 2e29); p["Position"][1] = (3, 4, 0)
    sim.init(p)
    acc = sim.accelerations()
    want = port.allpairs_forces(p) / p["Mass"][:, None]
    assert rel_err(acc, want).max() < 1e-5
    p = np.zeros(64, dtype=dt); p["Mass"] = 1e30
    p["Position"] = np.random.default_rng(0).uniform(5000, 9000, (64, 3)).astype(np.float32)
    sim.init(p)
    assert np.all(sim.accelerations() == 0)                  # no sources inside the root
    codes, _ = sim.morton()
    assert len(codes) == 0
    p = np.zeros(40, dtype=dt); p["Mass"] = 1e30
    p["Position"][:20] = (7.25, -3.5, 100.0)                 # 20 coincident bodies (the reference would recurse forever)
    p["Position"][20:] = np.random.default_rng(1).uniform(-500, 500, (20, 3)).astype(np.float32)
    sim.init(p)
    acc = sim.accelerations()
    want = port.allpairs_forces(p) / p["Mass"][:, None]      # small n: the tree walk opens everything near
    assert np.all(np.isfinite(acc))
    assert np.median(rel_err(acc, want)) < 2e-2
    sim.close()


def test_sharded_ranks_reproduce_single_gpu_bitwise(pkg):
    p = pkg.seed_galaxy_host(6000, 21, 1.0)
    one = bh(pkg, theta=0.5)
    one.init(p)
    acc = one.accelerations()
    shards = [bh(pkg, theta=0.5, rank=r, world=3) for r in range(3)]
    for s in shards:
        s.init(p)
    got = np.concatenate([s.accelerations() for s in shards])
    # same tree on every rank; a shard's warps group different targets, which changes which nodes
    # the WARP opens but neither what a lane accepts nor the order it accumulates in -> bit identical
    assert np.array_equal(got, acc)
    for s in shards:
        s.close()
    one.close()


def test_balanced_walk_over_peer_memory_reproduces_single_gpu_bitwise(pkg):
    """csrc/p2p.cu + tree.cu: with peer memory attached nb_step deals the Morton-ordered target list out
    to the ranks block by block and every rank stores the accelerations it computed straight into their
    owner's array (here: three shard handles on one GPU attached by handle), then the fused kick-drift
    pushes positions.  A lane's result does not depend on which rank or warp walks it, so after k steps
    the shards hold exactly what a single handle holds -- including bodies that left the root cube."""
    p = pkg.seed_galaxy_host(6000, 21, 1.0)
    p["Position"][100:110] *= 40.0                     # some bodies outside the cube: targets only
    dt, steps, world = 0.02 / 60, 7, 3
    one = bh(pkg, theta=0.5)
    one.init(p)
    one.step(dt, steps)
    want = p.copy()
    one.read(want)
    shards = [bh(pkg, theta=0.5, rank=r, world=world) for r in range(world)]
    for s in shards:
        s.init(p)
    for s in shards:
        s.p2p_attach_local(shards)
    for _ in range(steps):
        for s in shards:
            s.step(dt, 1)                              # one host thread drives all ranks, like one step of every process
    got = p.copy()
    for s in shards:
        s.read(got)                                    # each writes its owned range
    assert np.array_equal(got["Position"], want["Position"])
    assert np.array_equal(got["Velocity"], want["Velocity"])
    assert np.array_equal(got["Forces"], want["Forces"])          # BarnesHut leaves m*a of the last step
    acc = np.concatenate([s.accelerations() for s in shards])
    assert np.array_equal(acc, one.accelerations())
    # from the second step on the per-step sort is sharded by Morton-key range: every rank sorts one range
    # and stores its segment into every rank's sorted arrays -- the same codes in the same order everywhere
    codes, order = one.morton()
    for s in shards:
        c, o = s.morton()
        assert np.array_equal(c, codes) and np.array_equal(o, order)
    # a coarse step throws most arm bodies out of the cube: the key ranges of the previous step no longer
    # balance (one rank's range may even be empty) and the result must not care
    for _ in range(4):
        one.step(0.05, 1)
        for s in shards:
            s.step(0.05, 1)
    for s in shards:
        s.read(got)
    one.read(want)
    assert np.array_equal(got["Position"], want["Position"]) and np.array_equal(got["Velocity"], want["Velocity"])
    codes, order = one.morton()
    assert len(codes) < len(p) // 2                              # most bodies have left
    for s in shards:
        c, o = s.morton()
        assert np.array_equal(c, codes) and np.array_equal(o, order)
    for s in shards:
        s.close()
    one.close()


def test_large_n_sampled_parity(pkg):
    """configs[3] shape at N = 262144: sampled targets against the reference walk."""
    n = 1 << 18
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    sim = bh(pkg, theta=0.5)
    sim.init(p)
    acc = sim.accelerations()
    t = np.arange(0, n, n // 256)
    want = checker.barneshut_accel(p, 0.5, t)
    err = rel_err(acc[t], want)
    print("BH n=262144 rel err: median %.2e p99 %.2e" % (np.median(err), np.quantile(err, 0.99)))
    assert np.median(err) < BH_MEDIAN_RTOL
    # size-independent properties: sorted keys, a permutation, momentum balance
    codes, order = sim.morton()
    assert np.all(codes[1:] >= codes[:-1])
    assert np.array_equal(np.sort(order), np.arange(len(order)))
    sim.close()


def test_energy_drift_within_twice_the_reference(pkg):
    """BASELINE.json north_star: energy drift over 1000 steps within 2x the reference's.  Two-galaxy
    collision scene (config 5) at the size the reference's CPU path finishes in seconds; the
    reference's drift (tests/golden/energy_drift_n4096.npz, from oracle/_ref) and ours use the same
    estimator: E = sum 1/2 m v^2 + Scale * sum_{i<j} U(r), all pairs, fp64."""
    g = load_golden("energy_drift_n4096.npz")
    scene = as_particles(g["scene"], pkg.PARTICLE_DTYPE).copy()
    n = len(scene)
    mine = pkg.seed_collision_host(n, 42, 1.0, separation=float(g["separation"]), approach_speed=float(g["approach"]))
    for f in ("Position", "Velocity", "Mass"):
        assert np.array_equal(mine[f], scene[f])               # the product seeder builds the same scene
    sim = bh(pkg, theta=float(g["theta"]))
    sim.init(scene)
    ke0, pe0 = sim.energy()
    assert abs(ke0 - g["energies"][0, 0]) < 1e-9 * abs(ke0) and abs(pe0 - g["energies"][0, 1]) < 1e-6 * abs(pe0)
    e0 = ke0 + pe0
    drift = [0.0]
    for _ in range(10):
        sim.step(float(g["dt"]), 100)
        ke, pe = sim.energy()
        drift.append(abs(ke + pe - e0) / abs(e0))
    drift = np.array(drift)
    print("energy drift ours     :", drift)
    print("energy drift reference:", g["drift"])
    assert drift[-1] <= 2.0 * g["drift"][-1]
    assert drift.max() <= 2.0 * g["drift"].max()
    # and the trajectories are the reference's: the two drifts track each other
    assert np.all(np.abs(drift[1:] - g["drift"][1:]) <= 0.25 * g["drift"][1:] + 1e-6)
    sim.close()


def test_sampled_energy_estimator(pkg):
    """nb_energy_sampled: stride 1 is the exact pair sum of nb_energy (and of the oracle's
    port.energy); a stride-s sample is that sum restricted to bodies s*k, scaled by s."""
    g = load_golden("energy_drift_n4096.npz")
    scene = as_particles(g["scene"], pkg.PARTICLE_DTYPE).copy()
    n = len(scene)
    sim = bh(pkg, theta=0.5)
    sim.init(scene)
    ke, pe = sim.energy()
    ke1, pe1, ns1 = sim.energy_sampled(1)
    assert ns1 == n
    assert abs(ke1 - ke) <= 1e-12 * abs(ke) and abs(pe1 - pe) <= 1e-10 * abs(pe)
    want_ke, want_pe = port.energy(scene)
    assert abs(ke1 - want_ke) <= 1e-9 * abs(want_ke) and abs(pe1 - want_pe) <= 1e-6 * abs(want_pe)
    # stride 8 against a direct fp64 evaluation of the same 512 bodies x all sources
    stride = 8
    ke8, pe8, ns8 = sim.energy_sampled(stride)
    assert ns8 == n // stride and abs(ke8 - ke) <= 1e-12 * abs(ke)
    pos = scene["Position"].astype(np.float64)
    m = scene["Mass"]
    G, S, scale = 6.674e-11, 10.0, 2.3e13
    w = (G * m).astype(np.float32).astype(np.float64)          # the engine keeps G*m in fp32
    want = 0.0
    for i in range(0, n, stride):
        r = np.linalg.norm(pos - pos[i], axis=1)
        r[i] = np.inf
        want += 0.5 * m[i] / np.sqrt(S) * -(w * np.arctan(np.sqrt(S) / r)).sum()
    want *= stride * scale
    assert abs(pe8 - want) <= 1e-9 * abs(want)
    # and it estimates the whole: within the sampling noise of 512 bodies of a bimodal mass spectrum
    assert abs(pe8 - pe) < 0.25 * abs(pe)
    sim.close()


def test_energy_drift_at_65536_bodies_tracks_the_reference(pkg):
    """The same comparison at 16x the bodies (tests/golden/make_golden_drift64k.py: the reference's
    Barnes-Hut CPU path, 1000 steps, ~10 minutes on 4 workers): total mass and drift are several times
    larger than at 4096 bodies, so this is the sharper check that the GPU path integrates the same
    system.  Energies every 250 steps, exact pair sum on both sides."""
    g = load_golden("energy_drift_n65536.npz")
    n = int(g["n"])
    scene = pkg.seed_collision_host(n, 42, 1.0, separation=float(g["separation"]), approach_speed=float(g["approach"]))
    sim = bh(pkg, theta=float(g["theta"]))
    sim.init(scene)
    ke0, pe0 = sim.energy()
    assert abs(ke0 - g["energies"][0, 0]) < 1e-9 * abs(ke0) and abs(pe0 - g["energies"][0, 1]) < 1e-6 * abs(pe0)
    e0 = ke0 + pe0
    drift = [0.0]
    for _ in range(len(g["drift"]) - 1):
        sim.step(float(g["dt"]), int(g["every"]))
        ke, pe = sim.energy()
        drift.append(abs(ke + pe - e0) / abs(e0))
    drift = np.array(drift)
    print("energy drift ours      (N=65536):", drift)
    print("energy drift reference (N=65536):", g["drift"])
    assert drift[-1] <= 2.0 * g["drift"][-1] and drift.max() <= 2.0 * g["drift"].max()
    assert np.all(np.abs(drift[1:] - g["drift"][1:]) <= 0.25 * g["drift"][1:] + 1e-5)
    sim.close()


def test_leaf_cells_are_the_cubes_the_reference_draws(pkg):
    """nb_get_leaf_cells against BarnesHut::RenderDebug (Octree.cpp:147-175): one cube per occupied leaf,
    same bodies in the same (recursion = Morton) order, same centre and size.  The reference derives its
    bounds by repeated fp32 halving, exact for the levels these scenes reach."""
    for p in (pkg.seed_galaxy_host(1024, 42, 1.0), scattered(pkg, 3000, 5, duplicates=False), pkg.seed_galaxy_host(1, 3, 1.0)):
        sim = bh(pkg)
        sim.init(p)
        cells, body = sim.leaf_cells()
        if ref.available():
            want, want_body = ref.octree_leaf_cubes(p)
            assert np.array_equal(body, want_body)
            assert np.array_equal(cells[:, 3], want[:, 3])
            assert np.abs(cells[:, :3] - want[:, :3]).max() <= 1e-3
            exact = want[:, 3] >= 8000.0 / 2 ** 16
            assert np.array_equal(cells[exact], want[exact])
        depth, path = port.octree_paths(p)[:2]
        inb = depth[body] >= 0
        assert inb.all() and np.array_equal(cells[:, 3], (8000.0 / 2.0 ** depth[body]).astype(np.float32))
        inside = np.all(np.abs(p["Position"][body] - cells[:, :3]) <= cells[:, 3:4] / 2, axis=1)
        assert inside.all()                                   # every body lies in its cube
        # the export borrows the traversal-record buffer: the tree hooks and the next force pass are unaffected
        again, _ = sim.leaf_cells()
        assert np.array_equal(again, cells)
        assert np.all(np.isfinite(sim.accelerations()))
        sim.close()


def test_walk_occupancy_histogram_is_consistent_with_the_counters(pkg, galaxy):
    """nb_get_walk_occupancy: sum_k k * hist[k] is the number of node visits by lanes that needed them."""
    sim = bh(pkg, theta=0.5)
    sim.init(galaxy)
    st = sim.walk_stats()
    h = sim.walk_occupancy()
    assert int((h * np.arange(33, dtype=np.uint64)).sum()) == st["visits"]
    assert h[0] == 0 and h.sum() >= st["visits"] // 32
    sim.close()


@pytest.mark.parametrize("n", [4000, 50000])
def test_graph_replay_of_small_steps_is_bitwise_the_direct_launches(pkg, n):
    """One-GPU Barnes-Hut steps of small scenes replay their launches as CUDA graphs (csrc/nb_api.cu,
    compute_forces_graphed): same kernels and arguments, so the state after 20 steps must be bitwise the state the
    directly launched steps produce -- also across a theta change (the graphs are re-captured)."""
    p = pkg.seed_galaxy_host(n, 11, 1.0)
    hashes = []
    for graphs in (True, False):
        sim = bh(pkg, theta=0.5)
        sim.enable_graphs(graphs)
        sim.init(p)
        sim.step(0.02 / 60, 10)
        sim.set_theta(0.8)
        sim.step(0.02 / 60, 10)
        hashes.append(sim.state_hash())
        launches = sim.last_step_timing()[2]
        sim.close()
    assert hashes[0] == hashes[1]
    assert launches >= 100         # 10 steps x (13 kernels with the single-block sort at 4 000 bodies, ~40 at 50 000), counted either way


def test_alternative_build_kernels_give_the_same_tree_and_forces(pkg):
    """NB_BUILD=agglomerative (k_build_up: construction fused with the reduction, nodes numbered by split position)
    must produce the tree of the default two-pass build: same topology through nb_get_tree's renumbering, same node
    sums, bitwise the same accelerations.  Runs in a child process because the switch is read once per process."""
    import json, os, subprocess, sys
    from conftest import ROOT
    code = r"""
import importlib, json, sys, hashlib
import numpy as np
sys.path.insert(0, %r)
pkg = importlib.import_module("procedural-universe_b200")
p = pkg.seed_galaxy_host(30011, 5, 1.0)
sim = pkg.Sim(mode=pkg.MODE_BARNESHUT, theta=0.5)
sim.init(p)
t = sim.tree()
acc = sim.accelerations()
h = hashlib.sha256()
for k in ("left", "right", "prefix", "mass", "com"):
    h.update(np.ascontiguousarray(t[k]).tobytes())
h.update(acc.tobytes())
sim.step(0.02 / 60, 5)
print(json.dumps({"tree_and_acc": h.hexdigest(), "state": sim.state_hash()}))
""" % ROOT
    out = []
    # also the alternative sort (one histogram pass + decoupled look-back) and the global-atomic reduction
    for env in ({}, {"NB_BUILD": "agglomerative"}, {"NB_SORT": "onesweep"}, {"NB_REDUCE": "global"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert all(o == out[0] for o in out[1:])
