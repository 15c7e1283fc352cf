#!/bin/bash
# GPU round-2 run G (1 GPU): callers' resolutions (640x480, 1024x768), facade loop, Isaac shim image generation.
mkdir -p gpurun_out
T=r02g
B="python bench.py --steps 100 --warmup 20 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
timeout 300 $B --width 640 --height 480 > gpurun_out/${T}_bench_480p.json 2> gpurun_out/${T}_bench_480p.err
timeout 300 $B --width 1024 --height 768 > gpurun_out/${T}_bench_768p.json 2> gpurun_out/${T}_bench_768p.err
timeout 600 python tools/episode_bench.py --steps 500 --capsule > gpurun_out/${T}_facade_1080p.json 2> gpurun_out/${T}_facade_1080p.err
timeout 600 python tools/episode_bench.py --steps 500 --capsule --width 640 --height 480 > gpurun_out/${T}_facade_480p.json 2> gpurun_out/${T}_facade_480p.err
timeout 600 python tools/imagegen_bench.py > gpurun_out/${T}_imagegen.json 2> gpurun_out/${T}_imagegen.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r02g_facade*.json"))+sorted(glob.glob("gpurun_out/r02g_imagegen*.json")):
    try:
        print(f.split('/')[-1], open(f).read().strip().splitlines()[-1][:1500])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/${T}_*.err
