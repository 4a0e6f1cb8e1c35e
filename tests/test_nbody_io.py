"""`.nbody` particle files (reference SimulationState.cpp:229-277 reader, :317-331 writer) and the
closest-particle oracle.  CPU only: file I/O is host code, no kernel is involved."""
import os

import numpy as np
import pytest

from conftest import same_particles
from oracle import port, ref


def test_save_is_the_raw_particle_array(pkg, tmp_path):
    p = pkg.seed_galaxy_host(1000, 7, 1.0)
    path = tmp_path / "a.nbody"
    pkg.save_nbody(str(path), p)
    raw = path.read_bytes()
    assert len(raw) == 1000 * 104                       # no header, no count (file.write(&p, sizeof(p)) per body)
    assert raw == p.tobytes()


def test_load_round_trip_and_partial_record(pkg, tmp_path):
    p = pkg.seed_galaxy_host(257, 3, 1.0)
    path = tmp_path / "b.nbody"
    pkg.save_nbody(str(path), p)
    q = pkg.load_nbody(str(path), recentre=False)
    assert q.tobytes() == p.tobytes()                   # colours, forces and the padding travel too
    with open(path, "ab") as f:
        f.write(b"\x01" * 50)                           # a torn trailing record is dropped (`if (!infile) break`)
    q = pkg.load_nbody(str(path), recentre=False)
    assert len(q) == 257 and same_particles(q, p)


def test_recentring_matches_the_restatement(pkg, tmp_path):
    for n, seed in ((1, 1), (2, 5), (1000, 7), (4096, 42)):
        p = pkg.seed_galaxy_host(n, seed, 1.0)
        p["Position"] += np.float32(123.25)
        path = tmp_path / f"c{n}.nbody"
        pkg.save_nbody(str(path), p)
        got = pkg.load_nbody(str(path))
        want = port.recentre(p)
        assert same_particles(got, want), n
        # the mass-weighted centre of the result sits at the origin to fp32 rounding
        if n > 1:
            c = (got["Position"].astype(np.float64) * got["Mass"][:, None]).sum(axis=0) / got["Mass"].sum()
            assert np.all(np.abs(c) < 1e-3)


def test_recentring_of_massless_bodies_is_nan_like_the_reference(pkg):
    p = np.zeros(4, dtype=pkg.PARTICLE_DTYPE)            # TotalMass == 0 -> 0/0
    pkg.recentre(p)
    assert np.all(np.isnan(p["Position"]))


def test_missing_file_is_an_error(pkg, tmp_path):
    with pytest.raises(pkg.NBodyError, match="Could not read particle file"):
        pkg.load_nbody(str(tmp_path / "nope.nbody"))


def test_empty_file(pkg, tmp_path):
    path = tmp_path / "empty.nbody"
    path.write_bytes(b"")
    assert len(pkg.load_nbody(str(path), recentre=False)) == 0


# ---- closest particle: the reference's own test vectors pin the oracle ------------------------------

MATHS_TESTS_FIXTURE = [(0.0, 0.0, 0.0), (1.0, 0.0, 5.0), (2.0, 7.0, 0.0), (3.0, 16.0, 0.0), (60.0, 0.0, 10.0)]
MATHS_TESTS_CASES = [((1.0, 7.0, 1.0), 2), ((50.0, 2.0, 7.0), 4)]     # test/MathsTests.cpp:4-33


def fixture_particles(dtype):
    p = np.zeros(len(MATHS_TESTS_FIXTURE), dtype=dtype)
    p["Position"] = MATHS_TESTS_FIXTURE
    return p


def test_closest_particle_oracle_on_the_reference_test_vectors(pkg):
    p = fixture_particles(pkg.PARTICLE_DTYPE)
    for pos, want in MATHS_TESTS_CASES:
        assert port.closest_particle(p, pos) == want
        if ref.available():
            assert ref.closest_particle(p, pos) == want


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_closest_particle_restatement_equals_reference(pkg):
    rng = np.random.default_rng(5)
    p = pkg.seed_galaxy_host(5000, 11, 1.0)
    p["Position"][100] = p["Position"][40]               # an exact tie: the first index wins
    for k in range(50):
        pos = rng.uniform(-800, 800, 3).astype(np.float32)
        assert port.closest_particle(p, pos) == ref.closest_particle(p, pos)
    assert port.closest_particle(p, p["Position"][100]) == ref.closest_particle(p, p["Position"][100]) == 40
