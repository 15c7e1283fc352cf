#!/bin/bash
# GPU round-2 run S (8 GPUs): scaling 8 / 4 / 2 / 1 on one box; the N=8 line carries config 4 (8-scene batch: replicas vs tile rows).
mkdir -p gpurun_out
T=r02s
run() { # name gpus extra-flags...
  local name=$1 n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
}
Q="--steps 100 --warmup 20 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
run n8 8 --steps 100 --warmup 20 --no-cpu-baseline --no-collide --no-handoff
run n4 4 $Q
run n2 2 $Q
timeout 600 python bench.py $Q > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err
SAGE_PIPELINE_DEPTH=8 run n8_d8 8 $Q
SAGE_PIPELINE_DEPTH=4 run n8_d4 8 $Q
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02s_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "equal", d.get("multigpu_frame_equal"), d["e2e"].get("frame_equal"), "batch8", d.get("batch8"))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "lat", round(d["frame_latency_ms"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 500 gpurun_out/${T}_n8.err
