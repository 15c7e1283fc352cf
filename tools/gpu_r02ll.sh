#!/bin/bash
# GPU round-2 run LL (8 GPUs): the tile-row split after the composite work (E = 2 index walk on blocks): N = 8 bench line.
mkdir -p gpurun_out
T=r02ll
Q="--steps 100 --warmup 20 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 $Q > gpurun_out/${T}_n8.json 2> gpurun_out/${T}_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02ll_n8.json").read().strip().splitlines()[-1])
    print("n8 value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "equal", d.get("multigpu_frame_equal"), d["e2e"].get("frame_equal"))
    print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "lat", round(d["frame_latency_ms"],3))
except Exception as e:
    print("ERR", e)
PY
tail -c 300 gpurun_out/${T}_n8.err
