#!/bin/bash
# GPU round-2 run D: emit v3, non-persistent composite, carve-out, ballot ranking; variants; ncu; episode harness.
mkdir -p gpurun_out
T=r02d
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline"
timeout 300 $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for v in match s1 e2 c12 c8; do
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 300 $B > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
done
SAGE_NO_CARVEOUT=1 timeout 300 $B > gpurun_out/${T}_bench_nocarve.json 2> gpurun_out/${T}_bench_nocarve.err
SAGE_TILE_IPT=8 timeout 300 $B > gpurun_out/${T}_bench_tile8.json 2> gpurun_out/${T}_bench_tile8.err
SAGE_NO_GRAPH=1 timeout 300 $B > gpurun_out/${T}_bench_nograph.json 2> gpurun_out/${T}_bench_nograph.err
SAGE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 61 --launch-count 20 -f -o gpurun_out/${T}_frame_full python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_ncu.log 2>&1
timeout 600 python tools/run_benchmark_episode.py --steps 200 --out gpurun_out/${T}_episode.json > gpurun_out/${T}_episode.log 2>&1
tail -c 1500 gpurun_out/${T}_episode.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
