#!/bin/bash
# GPU round-2 run AA (1 GPU): k_emit offset look-back reading 32 / 128 / 256 predecessors per round trip.
mkdir -p gpurun_out
T=r02aa
timeout 900 python -m pytest tests/test_render_parity_gpu.py tests/test_render_spec_golden.py -m gpu -q --no-header -x 2>&1 | tail -2
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for v in lb1 lb4 lb8 lb1 lb4; do
  L=sage-3d_official_b200/libsage_splat_$v.so; [ $v = lb4 ] && L=sage-3d_official_b200/libsage_splat.so
  SAGE_SPLAT_LIB=$L timeout 300 $B > gpurun_out/${T}_bench_${v}_1080_$RANDOM.json 2>> gpurun_out/${T}_bench.err
done
for v in lb1 lb4; do
  L=sage-3d_official_b200/libsage_splat_$v.so; [ $v = lb4 ] && L=sage-3d_official_b200/libsage_splat.so
  SAGE_SPLAT_LIB=$L timeout 300 $B --width 640 --height 480 > gpurun_out/${T}_bench_${v}_480.json 2>> gpurun_out/${T}_bench.err
  SAGE_SPLAT_LIB=$L timeout 300 python tools/strip_rate.py --split 8 --ctx 6 > gpurun_out/${T}_strip8_$v.json 2>> gpurun_out/${T}_bench.err
  cat gpurun_out/${T}_strip8_$v.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02aa_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "emit", round(d["roofline"]["frame"]["stage_ms_serial"]["emit"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench.err
