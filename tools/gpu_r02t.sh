#!/bin/bash
# GPU round-2 run T (1 GPU): branch-free 4-entry batches in the composite loop (8 and 6 CTAs/SM) vs the per-lane-branch loop; parity of the variants.
mkdir -p gpurun_out
T=r02t
for v in b4c8 b4c6; do
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 900 python -m pytest tests/test_render_parity_gpu.py -m gpu -q --no-header -x 2>&1 | tail -3 > gpurun_out/${T}_pytest_$v.log
  tail -1 gpurun_out/${T}_pytest_$v.log
done
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "1920 1080" "640 480" "1024 768"; do
  set -- $res
  timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_base_$2.json 2> gpurun_out/${T}_bench_base_$2.err
  for v in b4c8 b4c6; do
    SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_${v}_$2.json 2> gpurun_out/${T}_bench_${v}_$2.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02t_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "composite", round(d["roofline"]["frame"]["stage_ms_serial"]["composite"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench_b4c8_1080.err
