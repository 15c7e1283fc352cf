#!/bin/bash
# GPU round-2 run DD (2 GPUs): the multi-GPU path after the composite rewrite - two-GPU tests, N=2 bench line (frame equality device + host),
# and one rank's block of an 8-way / 4-way split with the composite's trip width 4 (default there) and 2.
mkdir -p gpurun_out
T=r02dd
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --no-header 2>&1 | tail -3
Q="--steps 100 --warmup 20 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $Q > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err
for s in 8 4; do
  timeout 300 python tools/strip_rate.py --split $s --ctx 6 > gpurun_out/${T}_strip${s}_e4.json 2>> gpurun_out/${T}_strip.err
  SAGE_COMPOSITE_WIDE_MAX_CTAS=0 timeout 300 python tools/strip_rate.py --split $s --ctx 6 > gpurun_out/${T}_strip${s}_e2.json 2>> gpurun_out/${T}_strip.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02dd_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "equal", d.get("multigpu_frame_equal"), d["e2e"].get("frame_equal"))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "lat", round(d["frame_latency_ms"],3))
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r02dd_strip*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(round(v["device_fps"]),round(v["host_submit_us_per_frame"],1)) for k,v in d.items() if isinstance(v,dict)})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 400 gpurun_out/${T}_n2.err gpurun_out/${T}_strip.err
