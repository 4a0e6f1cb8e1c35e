"""nb_closest_particle (csrc/query.cu) against Maths::ClosestParticle: the reference's own test
vectors (test/MathsTests.cpp:4-33), the oracle on seeded galaxies, ties, non-finite positions and
the state after a step.  GPU only (pytest -m gpu); calls go through the C ABI."""
import numpy as np
import pytest

from oracle import checker, port, ref
from test_nbody_io import MATHS_TESTS_CASES, fixture_particles

pytestmark = pytest.mark.gpu


def oracle_closest(p, pos):
    return ref.closest_particle(p, pos) if ref.available() else port.closest_particle(p, pos)


def test_reference_test_vectors(pkg):
    p = fixture_particles(pkg.PARTICLE_DTYPE)            # Particle{{x, y, z}}: everything else zero, as in the reference test
    for mode in (pkg.MODE_ALLPAIRS, pkg.MODE_BARNESHUT):
        sim = pkg.Sim(mode=mode)
        sim.init(p)
        for pos, want in MATHS_TESTS_CASES:
            idx, d2 = sim.closest_particle(pos)
            assert idx == want
            assert d2 == np.float32(np.sum((np.float32(pos) - p["Position"][want]) ** 2))
        sim.close()


@pytest.mark.parametrize("n", [1, 31, 1000, 65537, 1 << 20])
def test_matches_oracle_on_galaxies(pkg, n):
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    sim = pkg.Sim(mode=pkg.MODE_ALLPAIRS)
    sim.init(p)
    rng = np.random.default_rng(n)
    queries = [rng.uniform(-900, 900, 3).astype(np.float32) for _ in range(12)]
    queries += [p["Position"][n // 2], p["Position"][0], p["Position"][n - 1], np.float32([1e30, 0, 0])]
    for q in queries:
        idx, _ = sim.closest_particle(q)
        assert idx == oracle_closest(p, q)
    sim.close()


def test_ties_pick_the_first_index_and_nonfinite_never_win(pkg):
    n = 5000
    p = pkg.seed_galaxy_host(n, 9, 1.0)
    p["Position"][4000] = p["Position"][123]
    p["Position"][77] = p["Position"][123]
    p["Position"][5] = (np.nan, 0, 0)
    p["Position"][6] = (np.inf, 0, 0)
    sim = pkg.Sim(mode=pkg.MODE_BARNESHUT)
    sim.init(p)
    idx, d2 = sim.closest_particle(p["Position"][123])
    assert idx == 77 == oracle_closest(p, p["Position"][123]) and d2 == 0.0
    # a query so far away that every squared distance overflows: nothing wins, id stays 0
    far = np.float32([3e38, 3e38, 3e38])
    assert sim.closest_particle(far)[0] == 0 == oracle_closest(p, far)
    sim.close()


def test_query_follows_the_simulation_state(pkg):
    p = pkg.seed_galaxy_host(4096, 42, 1.0)
    sim = pkg.Sim(mode=pkg.MODE_ALLPAIRS)
    sim.init(p)
    q = p.copy()
    for _ in range(3):
        sim.update(q, 0.01)                              # q now holds the positions the device holds
    for pos in ([0, 0, 0], [300, -200, 50], q["Position"][17]):
        assert sim.closest_particle(pos)[0] == oracle_closest(q, np.float32(pos))
    sim.close()
