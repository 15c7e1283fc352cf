#!/bin/bash
# GPU round-2 run A: parity tests, bench (graph on/off), ncu launch list + full capture of one frame.
mkdir -p gpurun_out
T=r02a
timeout 900 python -m pytest tests -m gpu -x -q --no-header 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log | tail -8
timeout 300 python bench.py --steps 300 --warmup 30 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.err
SAGE_NO_GRAPH=1 timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/${T}_bench_nograph.json 2> gpurun_out/${T}_bench_nograph.err
SAGE_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_launches.log 2>&1
SAGE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 43 --launch-count 14 -f -o gpurun_out/${T}_frame_full python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_ncu.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r02a_bench.json","gpurun_out/r02a_bench_nograph.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
