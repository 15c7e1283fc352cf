#!/bin/bash
# GPU round-2 run B: parity tests, bench (+ composite occupancy variants), ncu launch list + full capture.
mkdir -p gpurun_out
T=r02c
timeout 1200 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -40 > gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 300 --warmup 30 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.err
for v in c10 c8; do
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
done
SAGE_TILE_IPT=8 timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/${T}_bench_tile8.json 2> gpurun_out/${T}_bench_tile8.err
SAGE_DEPTH_IPT=16 timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/${T}_bench_depth16.json 2> gpurun_out/${T}_bench_depth16.err
SAGE_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${T}_launches.csv python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_launches.log 2>&1
SAGE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 61 --launch-count 20 -f -o gpurun_out/${T}_frame_full python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02c_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
