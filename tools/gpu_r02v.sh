#!/bin/bash
# GPU round-2 run V (1 GPU): composite CTAs/SM throttled at launch (SAGE_COMPOSITE_OCC) at three resolutions; pipeline depth at N=1;
# turnover of one rank's block of an 8-way split (host submit vs device).
mkdir -p gpurun_out
T=r02v
timeout 600 python -m pytest tests/test_render_parity_gpu.py -m gpu -q --no-header -x 2>&1 | tail -2
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "640 480" "1024 768" "1920 1080"; do
  set -- $res
  for occ in 0 3 4 5 6; do
    SAGE_COMPOSITE_OCC=$occ timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_occ${occ}_$2.json 2> gpurun_out/${T}_occ${occ}_$2.err
  done
done
for d in 2 4 5; do
  SAGE_PIPELINE_DEPTH=$d timeout 300 $B > gpurun_out/${T}_depth${d}_1080.json 2> gpurun_out/${T}_depth${d}_1080.err
done
for c in 3 6 8; do
  timeout 300 python tools/strip_rate.py --split 8 --ctx $c > gpurun_out/${T}_strip8_ctx$c.json 2> gpurun_out/${T}_strip8_ctx$c.err
  cat gpurun_out/${T}_strip8_ctx$c.json
done
timeout 300 python tools/strip_rate.py --split 4 --ctx 6 > gpurun_out/${T}_strip4_ctx6.json 2> gpurun_out/${T}_strip4_ctx6.err
cat gpurun_out/${T}_strip4_ctx6.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02v_*_*.json")):
    if "strip" in f: continue
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "composite", round(d["roofline"]["frame"]["stage_ms_serial"]["composite"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_occ4_480.err
