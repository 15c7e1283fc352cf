#!/bin/bash
# GPU round-2 run U (1 GPU): grid-stride k_project / k_sh_color + host-side caching: parity and bench.
mkdir -p gpurun_out
T=r02u
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -4 > gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
timeout 300 $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 $B --width 640 --height 480 > gpurun_out/${T}_bench_480.json 2> gpurun_out/${T}_bench_480.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02u_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench.err
