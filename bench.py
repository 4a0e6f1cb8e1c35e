#!/usr/bin/env python
"""Benchmark of the N-body hot path (BASELINE.json: body interactions/s and steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--secondary auto|none|a,b,..] [--impl reference]

One "step" is one INBodySim::Update on device-resident state: force pass + fused kick-drift (for N > 1 the
kick-drift kernel stores the new positions straight into every rank's position array over NVLink peer
memory; `--exchange nccl` selects an ncclAllGather instead).  Workloads:

    allpairs_1m   all-pairs fp32, N = 2^20, one spiral galaxy (GalaxySeeder seed 42)   [headline;
                  BASELINE.json configs[1]; with --gpus N the same bodies are sharded over N ranks]
    allpairs_16m  all-pairs fp32, N = 2^24 sharded over the ranks (configs[2])
    bh_1m/bh_16m  Barnes-Hut theta = 0.5, per-step rebuild, dt = 0.02/60 (configs[3])
    collision_64m two-galaxy collision, N = 2^26, Barnes-Hut theta = 0.5 (configs[4]); adds an "energy"
                  object with the drift of the conserved quantity (collision_1m: same scene, 2^20;
                  *_norm: body masses x 4096/N, the scene of tests/golden/energy_drift_n4096.npz)

Prints ONE JSON line (rank 0): the headline workload, and -- `--secondary auto`, the default -- the other
BASELINE configs that fit the GPUs at hand under "secondary" (each with its own ms_per_step, roofline,
parity, clocks, energy), plus a bitwise comparison of the multi-process path with a one-GPU rerun.
`--impl reference` times the reference's own CPU path (oracle/_ref, else the C port) on the host cores on a
bounded sample of the same workload; that arm never loads the product library.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "allpairs_1m": dict(mode="allpairs", n=1 << 20, dt=0.01, seed=42,
                        desc="all-pairs fp32 N=1048576 single spiral galaxy (GalaxySeeder seed 42), kick-drift dt=0.01"),
    "allpairs_256k": dict(mode="allpairs", n=1 << 18, dt=0.01, seed=42,
                          desc="all-pairs fp32 N=262144 single spiral galaxy (GalaxySeeder seed 42), kick-drift dt=0.01"),
    "allpairs_16m": dict(mode="allpairs", n=1 << 24, dt=0.01, seed=42, golden="bh_bh16m_sampled.npz",
                         desc="all-pairs fp32 N=16777216 single spiral galaxy sharded over ranks, positions exchanged every step"),
    "bh_4k": dict(mode="bh", n=4000, dt=0.02 / 60, seed=42, theta=0.5,
                  desc="Barnes-Hut theta=0.5 N=4000 (the low end of the reference's interactive range; its default is 1000, SimulationState.hpp:69), per-step LBVH rebuild, dt=0.02/60"),
    "bh_50k": dict(mode="bh", n=50000, dt=0.02 / 60, seed=42, theta=0.5,
                   desc="Barnes-Hut theta=0.5 N=50000 (the reference UI's particle cap), per-step LBVH rebuild, dt=0.02/60"),
    "bh_1m": dict(mode="bh", n=1 << 20, dt=0.02 / 60, seed=42, theta=0.5,
                  desc="Barnes-Hut theta=0.5 N=1048576 single spiral galaxy, per-step LBVH rebuild, dt=0.02/60"),
    "bh_16m": dict(mode="bh", n=1 << 24, dt=0.02 / 60, seed=42, theta=0.5, golden="bh_bh16m_sampled.npz",
                   desc="Barnes-Hut theta=0.5 N=16777216 single spiral galaxy, per-step LBVH rebuild, dt=0.02/60"),
    # BASELINE.json configs[4]: two-galaxy collision (seeds 42 / 43, centres 2000 apart, approaching at
    # 2e16), Barnes-Hut theta = 0.5, dt = 0.02/60.  The energy of the conserved quantity is estimated
    # before the first and after the last step from a fixed sample of bodies.
    "collision_64m": dict(mode="bh", n=1 << 26, dt=0.02 / 60, seed=42, theta=0.5, scene="collision", energy_stride=16384,
                          golden="bh_collision64m_sampled.npz",
                          desc="two-galaxy collision N=67108864 (GalaxySeeder seeds 42/43), Barnes-Hut theta=0.5, per-step LBVH rebuild, dt=0.02/60"),
    "collision_64m_norm": dict(mode="bh", n=1 << 26, dt=0.02 / 60, seed=42, theta=0.5, scene="collision", energy_stride=16384,
                               mass_scale=4096.0 / (1 << 26),
                               desc="two-galaxy collision N=67108864, body masses x 4096/N (total mass of the 4096-body reference scene), Barnes-Hut theta=0.5, dt=0.02/60"),
    "collision_1m": dict(mode="bh", n=1 << 20, dt=0.02 / 60, seed=42, theta=0.5, scene="collision", energy_stride=256,
                         golden="bh_collision1m_sampled.npz",
                         desc="two-galaxy collision N=1048576 (GalaxySeeder seeds 42/43), Barnes-Hut theta=0.5, per-step LBVH rebuild, dt=0.02/60"),
    "collision_1m_norm": dict(mode="bh", n=1 << 20, dt=0.02 / 60, seed=42, theta=0.5, scene="collision", energy_stride=256,
                              mass_scale=4096.0 / (1 << 20),
                              desc="two-galaxy collision N=1048576, body masses x 4096/N, Barnes-Hut theta=0.5, dt=0.02/60"),
}
COLLISION = dict(separation=2000.0, approach_speed=2e16)     # the scene of tests/golden/energy_drift_n4096.npz

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
# `ncu --set full` captures (1 GPU); only quoted for the exact workload they were taken on.
NCU_TRAFFIC = {
    ("allpairs_1m", 1): (48.300288e6 + 281.225984e6, "profiles/r2_allpairs_ncu_full.txt"),
    ("bh_16m", 1): (2.576252e9 + 1.016920e9, "profiles/r2_walk_ncu_full.txt (first steps, all 2^24 bodies inside the root cube)"),
}

FLOPS_PER_INTERACTION = 20   # 3 sub, 5 d^2, 1 add S, 1 sqrt, 1 div, 3 div, 3 mul, 3 add
# tree build, algorithmic bytes per body: Morton 28, sort 8 x 24, Karras 72, reduction 128
BUILD_BYTES_PER_BODY = 28 + 192 + 72 + 128


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the timed region: samples before it are dropped."""
        self.t0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for stamp, ln in self.lines:
            if stamp < getattr(self, "t0", 0.0) or stamp > t1:
                continue
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own path on the host cores (oracle/_ref, else the C port).  Nothing here
# touches the product library.
# ---------------------------------------------------------------------------------------------------
def seed_reference(wl):
    """The workload's bodies through the REFERENCE's seeder (oracle/_ref).  Without it (no prebuilt library)
    a numpy disk of the same extent stands in: the all-pairs rate does not depend on where the bodies are."""
    from oracle import checker, ref
    n = wl["n"]
    if ref.available():
        p = checker.seed_scene(n, wl.get("scene", "galaxy"), wl["seed"])
    else:
        rng = np.random.RandomState(wl["seed"])
        p = np.zeros(n, dtype=ref.PARTICLE_DTYPE)
        p["Position"][:, :2] = rng.uniform(-720.0, 720.0, size=(n, 2)).astype(np.float32)
        p["Position"][:, 2] = rng.normal(0.0, 16.0, size=n).astype(np.float32)
        p["Mass"] = rng.uniform(1e28, 1e30, size=n)
    if "mass_scale" in wl:
        p["Mass"] *= wl["mass_scale"]
    return p


def cpu_sample(p, mode, seconds, theta=0.5, dt=0.02 / 60):
    """Times the reference CPU path on a bounded sample of the workload.
    Returns (interactions/s, cores, description, kind)."""
    from oracle import port, ref
    n = len(p)
    if mode == "allpairs":
        if ref.available():
            w = max(1, min(ref.lib().ref_hardware_workers(), 512))     # the reference pool's own size: hardware_concurrency() - 1
            count = w * 2
            secs, used, _ = ref.bruteforce_block(p, 0, count, workers=w)       # calibration + warm-up
            rate = count * (n - 1) / secs
            count = int(max(1, (seconds * rate / (n - 1)) // used)) * used
            count = min(count, (n // used) * used)
            secs, used, _ = ref.bruteforce_block(p, 0, count, workers=used)
            return count * (n - 1) / secs, used, f"BruteForceCPU::Exec on targets [0,{count}) x {n} sources, {used} pool workers, {secs:.2f} s", "reference"
        count = 4
        t0 = time.perf_counter(); port.allpairs_forces(p, 0, count); secs = time.perf_counter() - t0
        count = int(max(4, seconds * count / secs))
        t0 = time.perf_counter(); port.allpairs_forces(p, 0, count); secs = time.perf_counter() - t0
        return count * (n - 1) / secs, 1, f"C port of BruteForceCPU::Exec on targets [0,{count}) x {n} sources, 1 thread, {secs:.2f} s", "port"
    # Barnes-Hut: ONE whole BarnesHut::Update (serial octree build + CalculateForce of every body on the
    # reference's thread pool + integrator), at N <= 2^20.  Larger workloads are timed on their first 2^20
    # bodies and the per-body cost is scaled by log2(N)/20 (flagged as extrapolated).
    n_full = n
    if n > (1 << 20):
        p = p[: 1 << 20]
        n = 1 << 20
    scale = np.log2(n) / np.log2(n_full)
    note = "" if n == n_full else f"; measured at N={n}, per-body cost extrapolated to N={n_full} by N log N"
    if ref.available():
        work = ref.barneshut_work(p, np.arange(0, n, max(1, n // 512))[:512], theta)
        per_target = (work["cell_evals"] + work["leaf_evals"]) / 512.0     # interactions per body of the reference walk
        _, secs, used = ref.barneshut_run(p, np.float32(dt), 1, theta, workers=0)
        return per_target * n / secs * scale, used, (f"one BarnesHut::Update (serial Octree build + CalculateForce on {used} pool workers + integrator) "
                                                     f"at N={n}: {secs:.2f} s{note}"), "reference"
    targets = np.arange(0, n, max(1, n // 4096))
    t0 = time.perf_counter(); _, work = port.barneshut_forces(p, targets, theta, want_counters=True); secs = time.perf_counter() - t0
    per_target = (work["cell_evals"] + work["leaf_evals"]) / float(len(targets))
    return per_target * n / (secs * n / len(targets)) * scale, 1, f"C port: octree build + walk of {len(targets)} sampled targets, {secs:.2f} s, scaled{note}", "port"


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation on the host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = seed_reference(wl)
    rates, desc, cores, kind = [], "", 1, "port"
    per_step_seconds = 4.0
    for i in range(args.warmup + args.steps):
        rate, cores, desc, kind = cpu_sample(p, wl["mode"], per_step_seconds, wl.get("theta", 0.5), wl["dt"])
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    unit = "interactions/s"
    n = wl["n"]
    # Barnes-Hut: ~1000 interactions per body at theta = 0.5 (measured by the instrumented reference walk)
    ms_per_step = (n * (n - 1) / value if wl["mode"] == "allpairs" else n * 1000.0 / value) * 1e3
    line = {
        "impl": "reference", "metric": "body interactions/s",
        "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "steps_per_s": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 accumulate / f32 geometry", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "bodies": n, "dt": wl["dt"],
                   "note": "each step is a bounded sample of the workload; ms_per_step is the sample rate scaled to a whole step"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------
class Ctx:
    """What every workload of one bench process shares."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
        # NB_BENCH_ONE_DEVICE=1 (debugging the multi-process logic on a one-GPU box): every rank uses device 0 and
        # torch.distributed runs on gloo; the library's peer-memory path (CUDA IPC) is the same
        self.one_device = os.environ.get("NB_BENCH_ONE_DEVICE") == "1"
        if self.one_device:
            self.local_rank = 0
        torch.cuda.set_device(self.local_rank)
        # a flag wait that can never be satisfied should end the run in minutes, not after the library's default
        os.environ.setdefault("NB_P2P_TIMEOUT_MS", "120000")
        if self.world > 1:
            import datetime
            if self.one_device:
                dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=600))
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank), timeout=datetime.timedelta(seconds=600))
        self.cdev = "cpu" if self.one_device else "cuda"
        self.pkg = importlib.import_module("procedural-universe_b200")
        self.multi = importlib.import_module("procedural-universe_b200.multi") if self.world > 1 else None
        # an explicit stream: torch's default stream is the legacy stream (handle 0), for which nb_create would
        # make a non-blocking stream of its own -- the L2 flush and the timing events must sit on the stream
        # the library launches on
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
        # what the flush itself costs (it sits inside every timed region): reported beside each result
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            self.flush.zero_()
        ev[0].record(self.stream)
        for _ in range(10):
            self.flush.zero_()
        ev[1].record(self.stream)
        torch.cuda.synchronize()
        self.flush_ms = ev[0].elapsed_time(ev[1]) / 10.0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="sum"):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.cdev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        box = [None] * self.world
        self.dist.all_gather_object(box, obj)
        return box

    def new_sim(self, wl, world=None, rank=None, splits=None):
        pkg = self.pkg
        world = self.world if world is None else world
        rank = self.rank if rank is None else rank
        mode = pkg.MODE_ALLPAIRS if wl["mode"] == "allpairs" else pkg.MODE_BARNESHUT
        return pkg.Sim(mode=mode, theta=wl.get("theta", 2.0), device=self.local_rank, rank=rank, world=world,
                       stream=self.stream.cuda_stream, source_splits=self.args.splits if splits is None else splits,
                       kernel_variant=self.args.variant)

    def connect(self, sim):
        if self.world > 1:
            if self.args.exchange == "p2p":
                self.multi.connect_p2p(sim, self.rank, self.world)   # CUDA IPC handles travel over torch.distributed
            else:
                self.multi.connect(sim, self.rank)                   # library-owned NCCL communicator


def seed_workload(pkg, wl, out=None):
    """The workload's bodies through the product's bit-exact host seeders (into `out` if given)."""
    n = wl["n"]
    if wl.get("scene") == "collision":
        p = pkg.seed_collision_host(n, wl["seed"], 1.0, **COLLISION)
    else:
        p = pkg.seed_galaxy_host(n, wl["seed"], 1.0)
    if "mass_scale" in wl:
        p["Mass"] *= wl["mass_scale"]
    if out is None:
        return p
    out[:] = p
    return out


T_START = time.perf_counter()


def log(ctx, msg):
    """Progress on stderr (every rank): a hung multi-process run must show where it stopped."""
    print(f"[bench {time.perf_counter() - T_START:7.1f}s rank {ctx.rank}] {msg}", file=sys.stderr, flush=True)


def guarded(fn, what):
    """Rank-0-only extras (checker runs, probes) must not throw: an exception on one rank alone would leave the
    other ranks inside the next collective."""
    try:
        return fn()
    except Exception as exc:
        return {"error": f"{what}: {type(exc).__name__}: {exc}"}


def _rel(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


def _stats(err):
    return {"median_rel_err": float(np.median(err)), "max_rel_err": float(np.max(err)), "targets": int(len(err))}


def sample_targets(n):
    return np.unique(np.linspace(0, n - 1, 64).astype(np.int64)).astype(np.uint32)


def parity_check(ctx, sim, wl, particles, after_first_step=False, direct=None):
    """Untimed.  Every rank evaluates the accelerations of the sampled targets it owns, at the INITIAL positions:
    before the first step (nb_get_accel_of: one extra force pass), or -- all-pairs above 2^20 bodies, where that
    pass costs as much as a step -- right after the first warm-up step from the accelerations that step kicked
    with (nb_get_step_accel_of), `direct` having been evaluated before it.
      golden:        Octree::CalculateForce / BruteForceCPU::Exec of the REFERENCE at this workload's full size,
                     generated once by tests/golden/make_golden_bh16m.py (needs the reference tree and 12-45 GB);
      live_checker:  the oracle run here (N <= 2^20);
      device_direct: the reference's all-pairs law restated operation by operation on the device
                     (nb_direct_accel: fp32 geometry without fusion, fp64 force and sum) -- usable at any N.
    Barnes-Hut at theta = 0.5 differs from a direct sum by ~1 % (median) in the reference itself."""
    n = wl["n"]
    out = {}

    def fast(targets):
        acc = sim.step_accelerations_of(targets) if after_first_step else sim.accelerations_of(targets)   # NaN rows: bodies of other ranks
        mine = np.isfinite(acc).all(axis=1)
        acc = np.where(mine[:, None], acc, 0.0)
        got = np.array(ctx.reduce(acc.reshape(-1))).reshape(-1, 3)
        return got

    gold_file = wl.get("golden")
    gold_path = os.path.join(ROOT, "tests", "golden", gold_file) if gold_file else None
    if gold_path and os.path.exists(gold_path):
        g = np.load(gold_path)
        targets = g["targets"]
        rec = np.ascontiguousarray(g["records"]).view(ctx.pkg.PARTICLE_DTYPE).reshape(-1)
        # the bodies themselves: the host array, or -- seeded on the device -- the device image of the sampled records
        mine = particles[targets] if particles is not None else (sim.aos_records(targets) if not after_first_step else None)
        same = None if mine is None else bool(all(np.array_equal(mine[f], rec[f]) for f in ("Position", "Velocity", "Mass", "Colour")))
        got = fast(targets)
        mass = rec["Mass"][:, None]
        entry = {"file": "tests/golden/" + gold_file, "same_bodies_as_reference_seeder": same}
        nd = len(g["direct_forces"])
        want_direct = g["direct_forces"] / mass[:nd]
        if wl["mode"] == "allpairs":
            entry["vs_reference_BruteForceCPU"] = _stats(_rel(got[:nd], want_direct))
        else:
            want = g["forces"] / mass
            finite = np.isfinite(want).all(axis=1)
            if finite.any():
                entry["vs_reference_Octree_CalculateForce"] = _stats(_rel(got[finite], want[finite]))
            else:
                entry["vs_reference_Octree_CalculateForce"] = None
                entry["note"] = ("the reference octree returns non-finite forces for all sampled targets at this size "
                                 "(fp32 centre-of-mass accumulation overflows, Octree.cpp:86-105); compared with its direct sum instead")
            entry["vs_reference_BruteForceCPU"] = _stats(_rel(got[:nd], want_direct))
            entry["reference_tree_vs_its_own_direct_sum"] = (_stats(_rel((g["forces"] / mass)[:nd][finite[:nd]], want_direct[finite[:nd]]))
                                                             if finite[:nd].any() else None)
        out["golden"] = entry

    tsel = sample_targets(n)
    got = fast(tsel)
    def device_direct():
        d = direct if direct is not None else sim.direct_accelerations(tsel)
        return dict(_stats(_rel(got, d)), what="fast path vs nb_direct_accel (restated reference law, fp64 sums) on the device")

    def live_checker():
        from oracle import checker
        t16 = tsel[:: max(1, len(tsel) // 16)][:16]
        want = checker.allpairs_accel(particles, t16) if wl["mode"] == "allpairs" else checker.barneshut_accel(particles, wl.get("theta", 0.5), t16)
        idx = np.searchsorted(tsel, t16)
        return dict(_stats(_rel(got[idx], want)), checker=checker.kind())

    if ctx.rank == 0:
        out["device_direct"] = guarded(device_direct, "nb_direct_accel")
        if particles is not None and n <= (1 << 20):
            out["live_checker"] = guarded(live_checker, "oracle")
    return out if ctx.rank == 0 else None


def state_hashes(ctx, sim):
    """Checksums of the device state on every rank: all ranks must hold bitwise equal positions; the velocity
    checksums of the shards add up (mod 2^64) to the checksum a single handle would report."""
    hp, hv = sim.state_hash()
    box = ctx.gather((hp, hv))
    return {"positions_equal_on_all_ranks": len({b[0] for b in box}) == 1, "positions": box[0][0],
            "velocities": sum(b[1] for b in box) % (1 << 64)}


def run_workload(ctx, name, steps, warmup, headline):
    """One workload on all ranks; returns the result object on rank 0."""
    torch, pkg, args = ctx.torch, ctx.pkg, ctx.args
    rank, world = ctx.rank, ctx.world
    wl = WORKLOADS[name]
    if args.energy_stride > 0 and "energy_stride" in wl:
        wl = dict(wl, energy_stride=args.energy_stride)
    n, dt = wl["n"], wl["dt"] * args.dt_scale
    e2e_steps = 0 if args.no_e2e or n * 104 > (2 << 30) else (steps if headline else min(steps, 5))
    if not headline and wl["mode"] == "allpairs" and n > (1 << 20):
        e2e_steps = 0                           # a 2^24-body all-pairs Update is 14 s on 8 GPUs: the device-resident steps are all that is run

    # pinned host AoS array: the caller's std::vector<Particle>
    t_seed = time.perf_counter()
    if e2e_steps:
        host = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
        particles = seed_workload(pkg, wl, host.numpy().view(pkg.PARTICLE_DTYPE))
        seeded_on = "host (nb_seed_host, pinned array)"
    else:
        particles = None                        # multi-GB scenes without an e2e leg: seeded on the device, below
        seeded_on = "device (nb_seed_*_device: parallel parse of the reference's random stream, bit-exact)"
    sim = ctx.new_sim(wl)
    if particles is not None:
        sim.init(particles)
    elif wl.get("scene") == "collision":
        sim.seed_collision_device(n, wl["seed"], 1.0, **COLLISION)
    else:
        sim.seed_galaxy_device(n, wl["seed"], 1.0)
    if particles is None and "mass_scale" in wl:
        sim.scale_masses(wl["mass_scale"])
    t_seed = time.perf_counter() - t_seed
    log(ctx, f"{name}: seeded on the {seeded_on.split()[0]} and initialised in {t_seed:.1f} s")
    ctx.connect(sim)
    first, count = sim.owned_range()
    log(ctx, f"{name}: initialised, peers connected")

    deferred = wl["mode"] == "allpairs" and n > (1 << 20)
    if deferred:
        direct = None
        if rank == 0:
            try:
                direct = sim.direct_accelerations(sample_targets(n))
            except Exception:
                direct = None
    else:
        parity = parity_check(ctx, sim, wl, particles)

    def total_energy():
        """(kinetic, potential estimate, samples) summed over the ranks; see nb_energy_sampled."""
        ke, pe, ns = sim.energy_sampled(wl["energy_stride"])
        ke, pe, ns = ctx.reduce([ke, pe, float(ns)])
        return ke, pe, int(ns)

    log(ctx, f"{name}: parity checked")
    energy = {"start": total_energy()} if "energy_stride" in wl else None

    def walk_counters():
        """One instrumented traversal (untimed) of this rank's targets, summed over the ranks."""
        mine = sim.walk_stats()
        inside = 0.0
        if rank == 0:
            try:
                inside = float(sim.inbounds())
            except Exception:
                inside = -1.0
        tw = ctx.reduce([mine["cell_evals"], mine["leaf_evals"], mine["visits"], inside])
        return {"cell_evals": int(tw[0]), "leaf_evals": int(tw[1]), "visits": int(tw[2]), "bodies_inside_root_cube": int(tw[3]), "rank0": mine}

    # ---- device-resident steps: nothing between the events but the L2 flush and nb_step ----------
    for w in range(warmup):
        ctx.flush.zero_()
        sim.step(dt, 1)
        if deferred and w == 0:
            parity = parity_check(ctx, sim, wl, particles, after_first_step=True, direct=direct)
    if not e2e_steps and n > (1 << 24):
        particles = None                        # multi-GB host image: the device holds the state from here on
    ctx.barrier()
    log(ctx, f"{name}: {warmup} warm-up steps done")
    walk_start = walk_counters() if wl["mode"] == "bh" else None
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    if sampler:
        sampler.mark()
    e0.record(ctx.stream)
    for _ in range(steps):
        ctx.flush.zero_()                   # evict L2 between timed steps
        sim.step(dt, 1)
    e1.record(ctx.stream)
    ctx.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = ctx.reduce([e0.elapsed_time(e1)], "max")[0]
    log(ctx, f"{name}: {steps} timed steps, {total_ms / steps:.3f} ms/step")
    # CUDA events recorded by the library on the launching stream around the dominant kernel and the tree build of
    # every step (a ring of 64), read only now
    kms, build_ms, timed_steps = sim.step_timing_mean(steps)
    period_ms = sim.step_period_mean(steps)[0] if steps > 1 else None
    launches = sim.last_step_timing()[2] * steps

    if energy is not None:
        energy["end"] = total_energy()
    hashes = state_hashes(ctx, sim) if world > 1 else None
    log(ctx, f"{name}: state hashed")

    walk = walk_counters() if wl["mode"] == "bh" else None       # the state the last timed steps ran on

    log(ctx, f"{name}: walk counted")
    # ---- end to end through the host-array contract ------------------------------------------
    e2e_s = None
    if e2e_steps:
        # the caller's array must hold the CURRENT state (the adapter's array always does: every Update writes it
        # back): a shard handle re-reads only its own records, and the tree code needs every rank to see the same bodies
        sim.read(particles)
        ctx.barrier()
        sim.update(particles, dt)           # warm-up of the AoS path
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sim.update(particles, dt)       # H2D of the caller's array, step, D2H write-back; synchronous
        ctx.barrier()
        e2e_s = ctx.reduce([time.perf_counter() - t0], "max")[0]

    log(ctx, f"{name}: counters, hashes, e2e done")
    fp32_peak = None
    if rank == 0:
        try:
            fp32_peak = sim.probe_fp32_peak()
        except Exception:
            fp32_peak = float("nan")
    sim.close()
    if rank != 0:
        return None

    unit = "interactions/s"
    if wl["mode"] == "allpairs":
        per_step = float(n) * float(n - 1)
    else:
        # accepted cells + direct pairs over all targets; the scene evolves (N = 2^24 bodies of these masses collapse
        # and disperse within tens of steps), so the count is taken before and after the timed steps and averaged
        per_step = 0.5 * float(walk_start["cell_evals"] + walk_start["leaf_evals"] + walk["cell_evals"] + walk["leaf_evals"])
    res = {
        "value": per_step * steps / (total_ms * 1e-3), "unit": unit, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "steps_per_s": steps / (total_ms * 1e-3), "bodies_per_s": float(n) * steps / (total_ms * 1e-3),
        "dtype": "f32" if wl["mode"] == "allpairs" else "f32 (f64 velocities / centre-of-mass)",
        "config": {"workload": wl["desc"], "name": name, "bodies": n, "dt": dt,
                   "parallelism": f"target-sharded x{world}" + (f", exchange={args.exchange}" if world > 1 else ""),
                   "l2": "256 MiB memset between timed steps (inside the timed region); sources (16 B/body) are re-read from L2 by design",
                   "l2_flush_ms_per_step": ctx.flush_ms,
                   "kernel_variant": args.variant, "seeded_on": seeded_on, "seed_and_init_s": t_seed},
        "gpu_launches": launches,
    }
    pk = peaks()
    nominal = 148 * 128 * 2 * (pk.get("sm_max_mhz", 1965.0) * 1e6) * 1e-12
    probe = fp32_peak * 1e-12
    if wl["mode"] == "allpairs":
        inter_per_launch = float(count) * float(n)          # this rank's launch, self term included
        kernel, evals = "k_allpairs_* (tiled all-pairs acceleration)", inter_per_launch
    else:
        mine = walk["rank0"]
        kernel, evals = "k_walk (warp-cooperative stackless traversal)", float(mine["cell_evals"] + mine["leaf_evals"])
    achieved = evals * FLOPS_PER_INTERACTION / (kms * 1e-3) * 1e-12
    res["roofline"] = {
        "bound": "fp32", "kernel": kernel, "achieved": achieved, "peak": nominal, "unit": "TFLOP/s", "frac": achieved / nominal,
        "traffic": None,
        "peak_source": "nominal FP32 FMA rate 148 SMs x 128 lanes x 2 flop x max SM clock (MEASURED_PEAKS.json has no FP32-pipe figure); "
                       "peak_probe = the same rate measured live by nb_probe_fp32_peak (pure FFMA/FFMA2 kernel)",
        "peak_probe": probe, "frac_of_probe": achieved / probe,
        "flops_per_interaction": FLOPS_PER_INTERACTION, "interactions_per_launch": evals,
        "kernel_ms": kms, "kernel_ms_steps_averaged": timed_steps, "kernel_share_of_step": kms * steps / total_ms,
        "step_ms_same_steps": period_ms,
    }
    if wl["mode"] == "bh":
        build_bytes = float(n) * BUILD_BYTES_PER_BODY
        hbm = pk.get("hbm_gbs", 6546.9)
        res["roofline"]["node_visits_per_launch"] = float(walk["rank0"]["visits"])
        res["roofline"]["note"] = "interactions_per_launch and kernel_ms both belong to the END of the run (last <= 64 steps)"
        res["interactions_per_step"] = {k: {"cells": w["cell_evals"], "pairs": w["leaf_evals"], "bodies_inside_root_cube": w["bodies_inside_root_cube"]}
                                        for k, w in (("before_timed_steps", walk_start), ("after_timed_steps", walk))}
        res["roofline"]["build"] = {"ms": build_ms, "bound": "hbm", "algorithmic_bytes": build_bytes,
                                    "achieved": build_bytes / (build_ms * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                                    "frac": build_bytes / (build_ms * 1e-3) * 1e-9 / hbm, "share_of_step": build_ms * steps / total_ms}
    if (name, world) in NCU_TRAFFIC and args.variant == 0:
        res["roofline"]["traffic"], res["roofline"]["traffic_source"] = NCU_TRAFFIC[(name, world)]
    if e2e_s is not None:
        res["e2e"] = {"value": per_step * e2e_steps / e2e_s, "unit": unit, "h2d_bytes_per_step": (n if world == 1 else count) * 104,
                      "d2h_bytes_per_step": count * 104, "ms_per_step": e2e_s * 1e3 / e2e_steps, "steps": e2e_steps,
                      "api": "nb_update_aos on a pinned 104-byte Particle array (INBodySim::Update contract)"}
    res["clocks"] = clocks
    res["parity"] = parity
    if hashes is not None:
        res["parity"]["multi_gpu_state"] = hashes
    if energy is not None:
        (k0, p0, ns), (k1, p1, _) = energy["start"], energy["end"]
        res["energy"] = {
            "steps": warmup + steps, "dt": dt, "physical_time": (warmup + steps) * dt, "kinetic": [k0, k1], "potential": [p0, p1],
            "drift": abs((k1 + p1) - (k0 + p0)) / abs(k0 + p0),
            "estimator": f"E = sum 1/2 m v^2 (exact) + 2.3e13 * sum U(r) estimated from {ns} bodies (every {wl['energy_stride']}th) x all sources, same bodies at both times",
        }
    if energy is not None and "mass_scale" in wl:
        # the scene with the total mass of the 4096-body golden: the reference's own Barnes-Hut path drifts by this much
        # over the same number of steps at the same dt (tests/golden/energy_drift_n4096.npz, checkpoints every 100 steps)
        try:
            g = np.load(os.path.join(ROOT, "tests", "golden", "energy_drift_n4096.npz"))
            k = min(len(g["drift"]) - 1, max(1, int(round((warmup + steps) / 100.0))))
            ref_drift = float(g["drift"][k])
            res["energy"]["reference_drift"] = {"n": 4096, "steps": 100 * k, "drift": ref_drift, "source": "tests/golden/energy_drift_n4096.npz (reference BarnesHut::Update)"}
            res["energy"]["ratio_vs_reference"] = res["energy"]["drift"] / ref_drift
        except Exception:
            pass
    return res


def run_bitwise(ctx, name, steps):
    """The multi-process path against one GPU: `steps` steps on all ranks (IPC peer memory, NVLink stores, step
    flags), then the same steps on rank 0 alone; the state checksums must be equal."""
    wl = WORKLOADS[name]
    particles = seed_workload(ctx.pkg, wl)
    # all-pairs: the number of source splits (0 = sized to fill the GPU for THIS rank's target count) decides the order
    # in which a target's partial sums are added -- pinned, so that N ranks and one GPU add in the same order
    splits = 8 if wl["mode"] == "allpairs" else None
    sim = ctx.new_sim(wl, splits=splits)
    sim.init(particles)
    ctx.connect(sim)
    sim.step(wl["dt"], steps)
    multi = state_hashes(ctx, sim)
    log(ctx, f"bitwise {name}: {steps} steps on {ctx.world} ranks hashed")
    ctx.barrier()
    sim.close()
    single = None
    if ctx.rank == 0:
        one = ctx.new_sim(wl, world=1, rank=0, splits=splits)
        one.init(particles)
        one.step(wl["dt"], steps)
        single = one.state_hash()
        one.close()
    ctx.barrier()
    if ctx.rank != 0:
        return None
    return {"steps": steps, "ranks": ctx.world, "positions_equal_on_all_ranks": multi["positions_equal_on_all_ranks"],
            "equals_single_gpu": bool(multi["positions_equal_on_all_ranks"] and multi["positions"] == single[0] and multi["velocities"] == single[1])}


def secondary_plan(args, world):
    """(workload, steps, warmup) beside the headline: the other BASELINE.json configs that fit the GPUs at hand.
    Barnes-Hut runs get >= 200 steps so that the 100 ms clock sampler sees the timed region."""
    if args.secondary == "none" or (args.secondary == "auto" and args.workload != "allpairs_1m"):
        return []
    if args.secondary != "auto":
        # an explicit list: Barnes-Hut scenes get enough steps for the 100 ms clock sampler (a 50 000-body step is 0.6 ms)
        long_runs = {"bh_50k": 2000, "bh_4k": 4000}
        return [(w, max(args.steps, long_runs.get(w, 200)) if WORKLOADS[w]["mode"] == "bh" else args.steps, 3)
                for w in args.secondary.split(",") if w]
    plan = [("bh_16m", 200, 5), ("bh_50k", 2000, 10)]
    if world == 1:
        plan.append(("bh_4k", 4000, 10))          # the reference's interactive range (UI.cpp:75 caps N at 50 000)
    if world >= 2:
        # configs[2]: 13.7 s per step on 8 GPUs, 110 s on 2 -- one timed step (two on 8 GPUs) after one untimed step
        plan.append(("allpairs_16m", 2 if world >= 8 else 1, 1))
    if world >= 8:
        plan.append(("collision_64m", 200, 5))
    return plan


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="allpairs_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default="auto", help="auto | none | comma-separated workloads run after the headline")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="all-pairs kernel table index")
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: fused kick-drift + peer-memory stores (p2p) or kick-drift + ncclAllGather (nccl)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--energy-stride", type=int, default=0, help="energy estimator: every k-th body x all sources (0 = the workload's default)")
    ap.add_argument("--dt-scale", type=float, default=1.0, help="multiplies the workload's dt (energy convergence runs: same physical time with --steps scaled up)")
    ap.add_argument("--bitwise-only", action="store_true", help="N > 1: only compare the multi-process path with a one-GPU rerun")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, wl)
        return

    ctx = Ctx(args)
    if args.bitwise_only:
        out = {name: run_bitwise(ctx, name, steps) for name, steps in (("bh_1m", 10), ("allpairs_256k", 5), ("collision_1m", 10))}
        if ctx.rank == 0:
            print(json.dumps({"multi_gpu_bitwise": bool(all(b["equals_single_gpu"] for b in out.values())), "detail": out}), flush=True)
        if ctx.world > 1:
            ctx.dist.destroy_process_group()
        return
    res = run_workload(ctx, args.workload, args.steps, args.warmup, headline=True)
    secondary, bitwise = {}, {}
    for name, steps, warmup in secondary_plan(args, ctx.world):
        try:
            r = run_workload(ctx, name, steps, warmup, headline=False)
        except Exception as exc:            # a secondary workload must not take the headline down with it
            r = {"error": f"{type(exc).__name__}: {exc}"}
            log(ctx, f"{name}: FAILED {r['error']}")
        if ctx.rank == 0:
            secondary[name] = r
    if ctx.world > 1 and args.secondary != "none" and args.exchange == "p2p":
        for name, steps in (("bh_1m", 10), ("allpairs_256k", 5)):
            try:
                b = run_bitwise(ctx, name, steps)
            except Exception as exc:
                b = {"error": f"{type(exc).__name__}: {exc}"}
            if ctx.rank == 0:
                bitwise[name] = b

    if ctx.rank == 0:
        line = {"metric": "body interactions/s", "value": res["value"], "unit": res["unit"], "n_gpus": ctx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "steps_per_s": res["steps_per_s"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic", "config": res["config"],
                "gpu_launches": res["gpu_launches"], "roofline": res["roofline"]}
        if "e2e" in res:
            line["e2e"] = res["e2e"]
        line["clocks"] = res["clocks"]
        line["parity"] = res["parity"]
        if "energy" in res:
            line["energy"] = res["energy"]
        if wl["mode"] == "bh":
            line["bodies_per_s"] = res["bodies_per_s"]
        if bitwise:
            line["parity"]["multi_gpu_bitwise"] = bool(all(b.get("equals_single_gpu") for b in bitwise.values()))
            line["parity"]["multi_gpu_bitwise_detail"] = bitwise
        if secondary:
            line["secondary"] = secondary
        if ctx.world == 1 and not args.no_cpu_baseline:
            p = seed_reference(wl)
            rate, cores, desc, kind = cpu_sample(p, wl["mode"], 12.0, wl.get("theta", 0.5), wl["dt"])
            line["cpu_baseline"] = {"value": rate, "unit": res["unit"], "cores": cores, "kind": kind, "sample": desc, "host_cores": host_cores()}
            # Barnes-Hut secondaries: one whole BarnesHut::Update of the reference on its thread pool, at the workload's
            # size up to 2^20 bodies; above that measured at 2^20 and flagged as extrapolated
            for name, r in secondary.items():
                swl = WORKLOADS[name]
                if swl["mode"] != "bh" or r is None or "error" in r:
                    continue
                p = seed_reference(dict(swl, n=min(swl["n"], 1 << 20)))
                rate, cores, desc, kind = cpu_sample(p, "bh", 0.0, swl.get("theta", 0.5), swl["dt"])
                scale = np.log2(len(p)) / np.log2(swl["n"])
                if len(p) != swl["n"]:
                    desc += f"; per-body cost extrapolated from N={len(p)} to N={swl['n']} by N log N"
                r["cpu_baseline"] = {"value": rate * scale, "unit": res["unit"], "cores": cores, "kind": kind, "sample": desc, "host_cores": host_cores()}
        print(json.dumps(line), flush=True)

    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
