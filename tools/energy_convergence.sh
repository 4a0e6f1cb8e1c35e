#!/bin/bash
# Energy evidence for BASELINE.json configs[4] (two-galaxy collision, 2^26 bodies, Barnes-Hut theta 0.5) on 8 GPUs:
#   1. dt-convergence: the same physical time (64 dt) with dt, dt/4, dt/16 -- a first-order symplectic integrator on an
#      under-resolved scene shows drift ~ dt; a kernel fault would not care about dt;
#   2. the 1000-step run of the config on the final code;
#   3. the same bodies with masses x 4096/N (total mass of the 4096-body scene the reference golden covers), 1000 steps.
# Writes one JSON line per run to $OUT (default gpurun_out/).
OUT=${OUT:-gpurun_out}
N=${N:-8}
run() { # name, workload, steps, dt-scale
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N \
     --workload $2 --steps $3 --warmup 3 --dt-scale $4 --energy-stride ${STRIDE:-2048} --secondary none --no-e2e --no-cpu-baseline \
     > $OUT/energy_$1.json 2> $OUT/energy_$1.err || tail -5 $OUT/energy_$1.err
}
run dt1 ${SCENE:-collision_64m} 61 1.0
run dt4 ${SCENE:-collision_64m} 253 0.25
run dt16 ${SCENE:-collision_64m} 1021 0.0625
run 1000steps ${SCENE:-collision_64m} 997 1.0
run norm_1000steps ${SCENE:-collision_64m}_norm 997 1.0
python - <<'PY'
import json, glob, os
out = os.environ.get("OUT", "gpurun_out")
for f in sorted(glob.glob(out + "/energy_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        e = d["energy"]
        print(os.path.basename(f), "steps", e["steps"], "dt", e["dt"], "drift", e["drift"], "ms/step", d["ms_per_step"], "ratio", e.get("ratio_vs_reference"))
    except Exception as exc:
        print(f, "unreadable", exc)
PY
