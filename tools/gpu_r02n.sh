#!/bin/bash
# GPU round-2 run N (1 GPU): 32x16 tiles (8 quadrant bits) vs 16x16 tiles, both with 7-bit tile-sort digits; parity first.
mkdir -p gpurun_out
T=r02n
timeout 1200 python -m pytest tests/test_render_parity_gpu.py -m gpu -q --no-header -x 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "1920 1080" "640 480" "1024 768"; do
  set -- $res
  timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_t32_$2.json 2> gpurun_out/${T}_bench_t32_$2.err
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_t16.so timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_t16_$2.json 2> gpurun_out/${T}_bench_t16_$2.err
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_c5.so timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_t32c5_$2.json 2> gpurun_out/${T}_bench_t32c5_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02n_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "D", int(d["config"]["mean_instances"]))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_*.err
