#!/bin/bash
# GPU round-2 run JJ (1 GPU): final evidence pass - per-stage box test on the E = 4 launches (an intermediate build of run HH still used it on frames of <= 4000 CTAs), conic deflated by the rounding
# allowance in every culling step, needle-splat tests: tests, bench (default + k20 + reference arm), resolutions, ncu launch list +
# full capture, facade.
mkdir -p gpurun_out
T=r02jj
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff > gpurun_out/${T}_bench_k20.json 2> gpurun_out/${T}_bench_k20.err
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
timeout 300 $B --width 640 --height 480 > gpurun_out/${T}_bench_480p.json 2> gpurun_out/${T}_bench_480p.err
timeout 300 $B --width 1024 --height 768 > gpurun_out/${T}_bench_768p.json 2> gpurun_out/${T}_bench_768p.err
SAGE_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${T}_launches.csv python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_launches.log 2>&1
SAGE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 61 --launch-count 20 -f -o gpurun_out/${T}_frame_full python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_ncu.log 2>&1
timeout 600 python tools/episode_bench.py --steps 500 --capsule > gpurun_out/${T}_facade_1080p.json 2> gpurun_out/${T}_facade_1080p.err
timeout 600 python tools/episode_bench.py --steps 500 --capsule --width 640 --height 480 > gpurun_out/${T}_facade_480p.json 2> gpurun_out/${T}_facade_480p.err
timeout 600 python tools/imagegen_bench.py > gpurun_out/${T}_imagegen.json 2> gpurun_out/${T}_imagegen.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02jj_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "roofline frac", round(d["roofline"]["frac"],4), "issue", (d["roofline"].get("issue") or {}).get("frac"), "clocks", d["clocks"])
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "V,D", round(d["config"]["mean_visible"]), round(d["config"]["mean_instances"]))
        for k in ("collide","batch8","cpu_baseline","handoff_jpeg"):
            if d.get(k): print("  ",k,str(d[k])[:300])
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r02jj_facade*.json"))+sorted(glob.glob("gpurun_out/r02jj_imagegen.json")):
    print(f.split('/')[-1], open(f).read().strip().splitlines()[-1][:420])
PY
tail -c 300 gpurun_out/${T}_bench.err gpurun_out/${T}_ncu.log
