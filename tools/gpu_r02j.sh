#!/bin/bash
# GPU round-2 run J (1 GPU): depth PNG tests, composite loop A/B (old vs diet) interleaved, PNG timing.
mkdir -p gpurun_out
T=r02j
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for rep in 1 2; do
  timeout 300 $B > gpurun_out/${T}_bench_new$rep.json 2> gpurun_out/${T}_bench_new$rep.err
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_oldcomp.so timeout 300 $B > gpurun_out/${T}_bench_old$rep.json 2> gpurun_out/${T}_bench_old$rep.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02j_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
PY
python tools/png_bench.py > gpurun_out/${T}_png.json 2> gpurun_out/${T}_png.err; cat gpurun_out/${T}_png.json
tail -c 400 gpurun_out/${T}_*.err
