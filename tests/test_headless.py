"""procedural-universe_b200/bin/nbody_headless: the reference's `nbody.exe --compute` run
(NBody.cpp:52-89 -> SimulationState::RunSimulation, SimulationState.cpp:279-332) and its RunBenchmark
protocol (:334-362) on the engine, over the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import checker

EXE = os.path.join(ROOT, "procedural-universe_b200", "bin", "nbody_headless")


def run(args, cwd):
    return subprocess.run([EXE] + args, cwd=cwd, capture_output=True, text=True, timeout=600)


def test_cli_is_built_and_keeps_the_log_format(tmp_path):
    assert os.path.exists(EXE), "run `make` (or __graft_entry__.build())"
    r = run([], tmp_path)
    assert r.returncode == 2
    assert re.fullmatch(r"\[Error\] .+\n", r.stdout)     # "[Error] text\n", reference test/LogTests.cpp:22-29
    r = run(["--bogus"], tmp_path)
    assert r.returncode == 2 and r.stdout.startswith("[Error] unknown option --bogus")


def test_without_a_gpu_the_run_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(["-c", "--steps", "1", "-p", "64"], tmp_path)
    assert r.returncode == 1 and "[Error] nb_create" in r.stdout
    assert not os.path.exists(tmp_path / "data")


@pytest.mark.gpu
def test_compute_run_writes_the_state_the_reference_would(pkg, tmp_path):
    n, steps = 2048, 5
    r = run(["-c", "--steps", str(steps), "-p", str(n), "-s", "0.6", "--sim", "allpairs", "--seeder", "galaxy", "--seed", "42",
             "--out", "state.nbody"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[Info] Brute Force (B200)\n" in r.stdout and "[Info] Wrote state.nbody\n" in r.stdout
    got = pkg.load_nbody(str(tmp_path / "state.nbody"), recentre=False)
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    dt = np.float32(0.6) * (np.float32(1.0) / np.float32(60.0))         # timestep * (1/60), NBody.cpp:87
    want = checker.allpairs_run(p, dt, steps)
    # accumulated kick against the reference's (bounds as in test_allpairs_gpu.test_ten_steps_n256_against_golden)
    kick = np.linalg.norm(want["Velocity"] - p["Velocity"], axis=1)
    err = np.linalg.norm(got["Velocity"] - want["Velocity"], axis=1) / kick
    assert np.median(err) < 1e-5 and err.max() < 1e-3
    assert np.abs(got["Position"] - want["Position"]).max() < 1e-2
    assert np.array_equal(got["Colour"], p["Colour"]) and np.all(got["Forces"] == 0)


@pytest.mark.gpu
def test_default_run_is_starsystem_barneshut_and_resumes_from_file(pkg, tmp_path):
    r = run(["-c", "--steps", "3", "-p", "500"], tmp_path)                 # reference defaults: StarSystem seeder, Barnes-Hut
    assert r.returncode == 0, r.stdout + r.stderr
    assert "[Info] Barnes-Hut (B200)\n" in r.stdout
    files = os.listdir(tmp_path / "data")
    assert len(files) == 1 and files[0].endswith(".nbody")
    first = pkg.load_nbody(str(tmp_path / "data" / files[0]), recentre=False)
    assert len(first) == 500 and first["Mass"][0] == 1e30                  # the star, StarSystemSeeder.cpp:18-55
    r = run(["-c", "--steps", "0", "-f", files[0], "--out", "again.nbody"], tmp_path)
    assert r.returncode == 0 and "[Info] Read 500 particles from file\n" in r.stdout
    again = pkg.load_nbody(str(tmp_path / "again.nbody"), recentre=False)
    want = pkg.load_nbody(str(tmp_path / "data" / files[0]), recentre=True)   # loaded state, recentred, NOT re-seeded
    assert np.array_equal(again["Position"], want["Position"]) and np.array_equal(again["Velocity"], want["Velocity"])


@pytest.mark.gpu
def test_benchmark_protocol(tmp_path):
    r = run(["--benchmark", "-p", "4000"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "[Info] Running benchmark" and lines[-1] == "[Info] Benchmark finished"
    got = [ln for ln in lines if ln.startswith("[Info] Benchmark ")]
    assert len(got) == 3                                   # one line per sim + "Benchmark finished"
    assert any("Brute Force (B200)" in ln and "ms/frame" in ln for ln in got)
    assert any("Barnes-Hut (B200)" in ln and "ms/frame" in ln for ln in got)
