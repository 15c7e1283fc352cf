#!/bin/bash
# GPU round-2 run II (1 GPU): the composite loop as a trip lambda with an index walk (no box test) and a candidate walk (box test):
# 1080p against the build of run CC/DD on the same box; parity.
mkdir -p gpurun_out
T=r02ii
timeout 900 python -m pytest tests/test_render_parity_gpu.py tests/test_render_spec_golden.py tests/test_env_facade.py -m gpu -q --no-header 2>&1 | tail -2
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for rep in 1 2; do
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_cc.so timeout 300 $B > gpurun_out/${T}_bench_cc_1080_$rep.json 2>> gpurun_out/${T}_bench.err
  timeout 300 $B > gpurun_out/${T}_bench_new_1080_$rep.json 2>> gpurun_out/${T}_bench.err
done
SAGE_COMPOSITE_PREFILTER=1 timeout 300 $B > gpurun_out/${T}_bench_newpre1_1080_1.json 2>> gpurun_out/${T}_bench.err
for res in "640 480" "1024 768"; do
  set -- $res
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_cc.so timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_cc_$2.json 2>> gpurun_out/${T}_bench.err
  timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_new_$2.json 2>> gpurun_out/${T}_bench.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ii_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "composite", round(d["roofline"]["frame"]["stage_ms_serial"]["composite"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench.err
