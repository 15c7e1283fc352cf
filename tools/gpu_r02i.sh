#!/bin/bash
# GPU round-2 run I (1 GPU): PDL along the front-half chain (end / early trigger / off), composite inner-loop diet.
mkdir -p gpurun_out
T=r02i
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
timeout 300 $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
SAGE_NO_PDL=1 timeout 300 $B > gpurun_out/${T}_bench_nopdl.json 2> gpurun_out/${T}_bench_nopdl.err
SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_pdlearly.so timeout 300 $B > gpurun_out/${T}_bench_pdlearly.json 2> gpurun_out/${T}_bench_pdlearly.err
timeout 300 $B --width 640 --height 480 > gpurun_out/${T}_bench_480p.json 2> gpurun_out/${T}_bench_480p.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 500 gpurun_out/${T}_*.err
