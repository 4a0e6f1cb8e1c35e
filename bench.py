#!/usr/bin/env python
"""Benchmark of the N-body hot path (BASELINE.json: body interactions/s and steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one INBodySim::Update on device-resident state: force pass + fused kick-drift
(+ the NCCL position all-gather when N > 1).  Workloads:

    allpairs_1m   all-pairs fp32, N = 2^20, one spiral galaxy (GalaxySeeder seed 42)   [default;
                  BASELINE.json configs[1]; with --gpus N the same bodies are sharded over N
                  ranks = strong scaling]
    allpairs_16m  all-pairs fp32, N = 2^24 sharded over the ranks (configs[2])
    bh_1m/bh_16m  Barnes-Hut theta = 0.5, per-step rebuild, dt = 0.02/60 (configs[3])
    collision_64m two-galaxy collision, N = 2^26, Barnes-Hut theta = 0.5 (configs[4]; run it with
                  --gpus 8 --steps 1000): adds an "energy" object with the drift of the conserved
                  quantity between the first and the last step (collision_1m: same scene, 2^20)

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU path
(oracle/_ref, else the C port) on the host cores on a bounded sample of the same workload.
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
    "allpairs_16m": dict(mode="allpairs", n=1 << 24, dt=0.01, seed=42,
                         desc="all-pairs fp32 N=16777216 single spiral galaxy sharded over ranks, NCCL position all-gather"),
    "bh_1m": dict(mode="bh", n=1 << 20, dt=0.02 / 60, seed=42, theta=0.5,
                  desc="Barnes-Hut theta=0.5 N=1048576 single spiral galaxy, per-step LBVH rebuild, dt=0.02/60"),
    "bh_16m": dict(mode="bh", n=1 << 24, dt=0.02 / 60, seed=42, theta=0.5,
                   desc="Barnes-Hut theta=0.5 N=16777216 single spiral galaxy, per-step LBVH rebuild, dt=0.02/60"),
    # BASELINE.json configs[4]: two-galaxy collision (seeds 42 / 43, centres 2000 apart, approaching at
    # 2e16), Barnes-Hut theta = 0.5, dt = 0.02/60; run with --steps 1000.  The energy of the conserved
    # quantity is estimated before the first and after the last step from a fixed sample of bodies.
    "collision_64m": dict(mode="bh", n=1 << 26, dt=0.02 / 60, seed=42, theta=0.5, scene="collision", energy_stride=16384,
                          desc="two-galaxy collision N=67108864 (GalaxySeeder seeds 42/43), Barnes-Hut theta=0.5, per-step LBVH rebuild, dt=0.02/60"),
    "collision_1m": dict(mode="bh", n=1 << 20, dt=0.02 / 60, seed=42, theta=0.5, scene="collision", energy_stride=256,
                         desc="two-galaxy collision N=1048576 (GalaxySeeder seeds 42/43), Barnes-Hut theta=0.5, per-step LBVH rebuild, dt=0.02/60"),
}
COLLISION = dict(separation=2000.0, approach_speed=2e16)     # the scene of tests/golden/energy_drift_n4096.npz

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
# `ncu --set full` captures (1 GPU); only quoted for the exact workload they were taken on.
NCU_TRAFFIC = {
    ("allpairs_1m", 1): (46.866432e6 + 281.826304e6, "profiles/r1_allpairs_fold2_ncu_full.txt"),
    ("bh_16m", 1): (2.639446e9 + 1.037880e9, "profiles/r1_bh_16m_ncu_full.txt"),
}

FLOPS_PER_INTERACTION = 20   # 3 sub, 5 d^2, 1 add S, 1 sqrt, 1 div, 3 div, 3 mul, 3 add


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(p, mode, seconds, theta=0.5):
    """Times the reference CPU path on a bounded block of targets x all sources.
    Returns (interactions/s or target-evals/s, cores, description, kind)."""
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
    # Barnes-Hut: one tree build + a sample of targets, scaled to a whole step.  Above 2^20 bodies the
    # reference octree (136 B x ~3.9 N nodes, built serially) is timed on the first 2^20 bodies and the
    # per-body cost is scaled by log2(N)/20 (flagged as extrapolated).
    n_full = n
    if n > (1 << 20):
        p = p[: 1 << 20]
        n = 1 << 20
    targets = np.arange(0, n, max(1, n // 4096))
    if ref.available():
        f, build, walk = ref.barneshut_forces(p, targets, theta)
        step = build + walk * n / len(targets)
        scale = np.log2(n) / np.log2(n_full)
        note = "" if n == n_full else f"; measured at N={n}, extrapolated to N={n_full} by N log N"
        work = ref.barneshut_work(p, targets[:512], theta)
        per_target = (work["cell_evals"] + work["leaf_evals"]) / 512.0     # interactions per body of the reference walk
        return per_target * n / step * scale, 1, (f"BarnesHut: Octree build {build:.2f} s (serial, as the reference) + CalculateForce on {len(targets)} sampled "
                                     f"targets {walk:.2f} s scaled to {n} targets on 1 thread{note}"), "reference"
    t0 = time.perf_counter(); _, work = port.barneshut_forces(p, targets, theta, want_counters=True); secs = time.perf_counter() - t0
    per_target = (work["cell_evals"] + work["leaf_evals"]) / float(len(targets))
    return per_target * n / (secs * n / len(targets)), 1, f"C port: octree build + walk of {len(targets)} sampled targets, {secs:.2f} s, scaled", "port"


def seed_workload(pkg, wl, out=None):
    """The workload's bodies through the product's bit-exact host seeders (into `out` if given)."""
    n = wl["n"]
    if wl.get("scene") == "collision":
        p = pkg.seed_collision_host(n, wl["seed"], 1.0, **COLLISION)
    else:
        p = pkg.seed_galaxy_host(n, wl["seed"], 1.0)
    if out is None:
        return p
    out[:] = p
    return out


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation on the host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = importlib.import_module("procedural-universe_b200")
    p = seed_workload(pkg, wl)
    rates, desc, cores, kind = [], "", 1, "port"
    per_step_seconds = 4.0
    for i in range(args.warmup + args.steps):
        rate, cores, desc, kind = cpu_sample(p, wl["mode"], per_step_seconds, wl.get("theta", 0.5))
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
        "config": {"workload": wl["desc"], "note": "each step is a bounded sample of the workload; ms_per_step is the sample rate scaled to a whole step"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="allpairs_1m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="all-pairs kernel table index")
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: fused kick-drift + peer-memory stores (p2p) or kick-drift + ncclAllGather (nccl)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pkg = importlib.import_module("procedural-universe_b200")
    n, dt = wl["n"], wl["dt"]
    mode = pkg.MODE_ALLPAIRS if wl["mode"] == "allpairs" else pkg.MODE_BARNESHUT

    # pinned host AoS array: the caller's std::vector<Particle>
    if n * 104 <= (2 << 30) and not args.no_e2e:
        host = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
        particles = seed_workload(pkg, wl, host.numpy().view(pkg.PARTICLE_DTYPE))
    else:
        particles = seed_workload(pkg, wl)      # multi-GB scenes: pageable, the e2e leg is not run on them
        args.no_e2e = True

    stream = torch.cuda.current_stream()
    sim = pkg.Sim(mode=mode, theta=wl.get("theta", 2.0), device=local_rank, rank=rank, world=world,
                  stream=stream.cuda_stream, source_splits=args.splits, kernel_variant=args.variant)
    sim.init(particles)
    if world > 1:
        multi = importlib.import_module("procedural-universe_b200.multi")
        if args.exchange == "p2p":
            multi.connect_p2p(sim, rank, world)   # CUDA IPC handles travel over torch.distributed
        else:
            multi.connect(sim, rank)           # library-owned NCCL communicator; id travels over torch.distributed
    first, count = sim.owned_range()
    if args.no_e2e and n > (1 << 24):
        particles = None                        # multi-GB host image: the device holds the state from here on

    # parity spot check against the oracle (untimed): a few owned targets x all N sources
    parity = None
    if rank == 0 and particles is not None:
        from oracle import checker
        tsel = first + np.arange(0, count, max(1, count // 16))[:16]
        acc = sim.accelerations()[tsel - first]
        want = None
        if wl["mode"] == "allpairs":
            want = checker.allpairs_accel(particles, tsel)
        elif n <= (1 << 20):      # the reference octree of 16 M bodies needs ~9 GB and minutes
            want = checker.barneshut_accel(particles, wl.get("theta", 0.5), tsel)
        if want is not None:
            err = np.linalg.norm(acc - want, axis=1) / np.linalg.norm(want, axis=1)
            parity = {"checker": checker.kind(), "targets": int(len(tsel)), "max_rel_err": float(err.max()),
                      "median_rel_err": float(np.median(err))}

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def total_energy():
        """(kinetic, potential estimate, samples) summed over the ranks; see nb_energy_sampled."""
        ke, pe, ns = sim.energy_sampled(wl["energy_stride"])
        t = torch.tensor([ke, pe, float(ns)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return float(t[0].item()), float(t[1].item()), int(t[2].item())

    energy = None
    if "energy_stride" in wl:
        energy = {"start": total_energy()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps --------------------------------------------------------------
    for _ in range(args.warmup):
        flush.zero_()
        sim.step(dt, 1)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = [], 0
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        flush.zero_()                       # evict L2 between timed steps
        sim.step(dt, 1)
        _, fms, k = sim.last_step_timing()  # CUDA events on the launching stream, inside the library
        kernel_ms.append(fms)
        launches += k
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    if energy is not None:
        energy["end"] = total_energy()

    walk, build_ms = None, None
    if wl["mode"] == "bh":
        build_ms = sim.last_build_ms()
        walk = sim.walk_stats()            # one instrumented traversal, untimed; this rank's targets
        if world > 1:
            tw = torch.tensor([walk["cell_evals"], walk["leaf_evals"], walk["visits"]], dtype=torch.float64, device="cuda")
            mine = dict(walk)
            dist.all_reduce(tw)
            walk = {"cell_evals": int(tw[0].item()), "leaf_evals": int(tw[1].item()), "visits": int(tw[2].item()),
                    "rank0": mine}

    # ---- end to end through the host-array contract ------------------------------------------
    e2e = None
    if not args.no_e2e:
        sim.update(particles, dt)           # warm-up of the AoS path
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sim.update(particles, dt)       # H2D of the caller's array, step, D2H write-back; synchronous
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        h2d = (n if world == 1 else count) * 104
        d2h = count * 104

    fp32_peak = sim.probe_fp32_peak() if rank == 0 else None

    if rank == 0:
        unit, metric = "interactions/s", "body interactions/s"
        if wl["mode"] == "allpairs":
            per_step = float(n) * float(n - 1)
        else:
            # the tree code's interactions: accepted cells + direct pairs, summed over all targets
            per_step = float(walk["cell_evals"] + walk["leaf_evals"])
        value = per_step * args.steps / (total_ms * 1e-3)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "steps_per_s": args.steps / (total_ms * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if wl["mode"] == "allpairs" else "f32 (f64 velocities / centre-of-mass)",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "bodies": n, "dt": dt, "parallelism": f"target-sharded x{world}" + (f", exchange={args.exchange}" if world > 1 else ""),
                       "l2": "256 MiB memset between timed steps (inside the timed region); sources (16 B/body) are re-read from L2 by design",
                       "kernel_variant": args.variant},
            "gpu_launches": launches,
        }
        pk = peaks()
        kms = float(np.mean(kernel_ms))
        if wl["mode"] == "allpairs":
            inter_per_launch = float(count) * float(n)          # this rank's launch, self term included
            achieved = inter_per_launch * FLOPS_PER_INTERACTION / (kms * 1e-3) * 1e-12
            nominal = 148 * 128 * 2 * (pk.get("sm_max_mhz", 1965.0) * 1e6) * 1e-12
            peak = fp32_peak * 1e-12
            line["roofline"] = {
                "bound": "fp32", "kernel": "k_allpairs_* (tiled all-pairs acceleration)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "measured live by nb_probe_fp32_peak (pure FFMA/FFMA2 issue-rate kernel on this GPU); MEASURED_PEAKS.json has no FP32-pipe figure",
                "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                "flops_per_interaction": FLOPS_PER_INTERACTION, "interactions_per_launch": inter_per_launch,
                "kernel_ms": kms, "kernel_share_of_step": kms * args.steps / total_ms,
            }
        else:
            mine = walk.get("rank0", walk)
            evals_launch = float(mine["cell_evals"] + mine["leaf_evals"])
            achieved = evals_launch * FLOPS_PER_INTERACTION / (kms * 1e-3) * 1e-12
            peak = fp32_peak * 1e-12
            # tree build: algorithmic bytes per body: Morton 28, sort 8 x 24, Karras 72, reduce 128
            build_bytes = float(n) * (28 + 192 + 72 + 128)
            line["bodies_per_s"] = float(n) * args.steps / (total_ms * 1e-3)
            line["roofline"] = {
                "bound": "fp32", "kernel": "k_walk (warp-cooperative stackless traversal)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "measured live by nb_probe_fp32_peak; the walk is issue-bound (28 instructions per node visit, ~50% of lane slots do an interaction), see DESIGN.md K7",
                "flops_per_interaction": FLOPS_PER_INTERACTION, "interactions_per_launch": evals_launch,
                "kernel_ms": kms, "kernel_share_of_step": kms * args.steps / total_ms,
                "node_visits_per_launch": float(mine["visits"]),
                "build": {"ms": build_ms, "bound": "hbm", "algorithmic_bytes": build_bytes,
                          "achieved": build_bytes / (build_ms * 1e-3) * 1e-9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                          "frac": build_bytes / (build_ms * 1e-3) * 1e-9 / pk.get("hbm_gbs", 6546.9)},
            }
        if (args.workload, world) in NCU_TRAFFIC and args.variant == 0:
            line["roofline"]["traffic"], line["roofline"]["traffic_source"] = NCU_TRAFFIC[(args.workload, world)]
        if e2e is None and not args.no_e2e:
            line["e2e"] = {"value": per_step * args.steps / e2e_s, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "ms_per_step": e2e_s * 1e3 / args.steps,
                           "api": "nb_update_aos on a pinned 104-byte Particle array (INBodySim::Update contract)"}
        line["clocks"] = clocks
        line["parity"] = parity
        if energy is not None:
            (k0, p0, ns), (k1, p1, _) = energy["start"], energy["end"]
            ref_drift = None
            try:
                ref_drift = float(np.load(os.path.join(ROOT, "tests", "golden", "energy_drift_n4096.npz"))["drift"][-1])
            except Exception:
                pass
            line["energy"] = {
                "steps": args.warmup + args.steps, "kinetic": [k0, k1], "potential": [p0, p1],
                "drift": abs((k1 + p1) - (k0 + p0)) / abs(k0 + p0),
                "estimator": f"E = sum 1/2 m v^2 (exact) + 2.3e13 * sum U(r) estimated from {ns} bodies (every {wl['energy_stride']}th) x all sources, same bodies at both times",
                "reference_drift_n4096_1000_steps": ref_drift,
            }
        if world == 1 and not args.no_cpu_baseline:
            p = seed_workload(pkg, wl)   # the initial state (particles now holds the state after the e2e steps)
            rate, cores, desc, kind = cpu_sample(p, wl["mode"], 12.0, wl.get("theta", 0.5))
            line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": cores, "kind": kind, "sample": desc,
                                    "host_cores": host_cores()}
        print(json.dumps(line), flush=True)

    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
