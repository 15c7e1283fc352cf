#!/bin/bash
# GPU round-2 run R (1 GPU): parity after the binary32 box + PDL on the JPEG/PNG chains; quick bench; JPEG / PNG timing.
mkdir -p gpurun_out
T=r02r
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide"
timeout 300 $B > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python tools/png_bench.py > gpurun_out/${T}_png.json 2> gpurun_out/${T}_png.err
timeout 300 python tools/jpeg_prof.py > gpurun_out/${T}_jpeg.log 2>&1
timeout 600 python tools/imagegen_bench.py > gpurun_out/${T}_imagegen.json 2> gpurun_out/${T}_imagegen.err
python - <<'PY'
import json,glob
d=json.loads(open("gpurun_out/r02r_bench.json").read().strip().splitlines()[-1])
print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "handoff", d.get("handoff_jpeg",{}).get("value"))
print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
print(open("gpurun_out/r02r_png.json").read()); print(open("gpurun_out/r02r_imagegen.json").read()[:600])
PY
tail -5 gpurun_out/${T}_jpeg.log
