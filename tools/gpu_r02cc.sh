#!/bin/bash
# GPU round-2 run CC (1 GPU): default bench.py after the hand-off leg fix; where the 640x480 facade step spends its time
# (pipeline depth 3 / 5, composite trip width 2 / 4).
mkdir -p gpurun_out
T=r02cc
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
E="python tools/episode_bench.py --steps 500 --capsule --width 640 --height 480"
timeout 300 $E > gpurun_out/${T}_facade_480p_d5_e4.json 2> gpurun_out/${T}_facade.err
SAGE_PIPELINE_DEPTH=3 timeout 300 $E > gpurun_out/${T}_facade_480p_d3_e4.json 2>> gpurun_out/${T}_facade.err
SAGE_COMPOSITE_WIDE_MAX_CTAS=0 timeout 300 $E > gpurun_out/${T}_facade_480p_d5_e2.json 2>> gpurun_out/${T}_facade.err
SAGE_PIPELINE_DEPTH=3 SAGE_COMPOSITE_WIDE_MAX_CTAS=0 timeout 300 $E > gpurun_out/${T}_facade_480p_d3_e2.json 2>> gpurun_out/${T}_facade.err
SAGE_PIPELINE_DEPTH=1 timeout 300 $E > gpurun_out/${T}_facade_480p_d1_e4.json 2>> gpurun_out/${T}_facade.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02cc_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "roofline frac", round(d["roofline"]["frac"],4), "issue", d["roofline"].get("issue",{}).get("frac"), "clocks", d["clocks"])
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
        for k in ("collide","batch8","cpu_baseline","handoff_jpeg"):
            if d.get(k): print("  ",k,d[k])
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r02cc_facade_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ("steps_per_s","ms_per_step","render_ms","collide_ms","capsule_ms","facade_perf","pipeline_depth")})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 400 gpurun_out/${T}_bench.err gpurun_out/${T}_facade.err
