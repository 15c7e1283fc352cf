#!/bin/bash
# GPU round-2 run E (1 GPU): tests, the new bench.py (all legs), episode harness at 1080p, ncu launch list.
mkdir -p gpurun_out
T=r02e
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 800 gpurun_out/${T}_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff > gpurun_out/${T}_bench_k20.json 2> gpurun_out/${T}_bench_k20.err
timeout 600 python tools/run_benchmark_episode.py --steps 300 --out gpurun_out/${T}_episode.json > gpurun_out/${T}_episode.log 2>&1
timeout 600 python tools/run_benchmark_episode.py --steps 300 --jpeg-handoff --out gpurun_out/${T}_episode_handoff.json > gpurun_out/${T}_episode_handoff.log 2>&1
SAGE_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${T}_launches.csv python tools/frame_prof.py 2000000 6 > gpurun_out/${T}_launches.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02e_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3),"blocks",d["timed_blocks"]["blocks"],d["timed_blocks"]["frames_per_s_min_max"],"clocks",d["clocks"])
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "graph", d["graph"])
        print("  collide",d.get("collide")); print("  batch8",d.get("batch8")); print("  cpu",d.get("cpu_baseline")); print("  handoff", d.get("handoff_jpeg"))
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r02e_episode*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], {k:d[k] for k in ("steps","steps_per_s","ms_per_step","render_ms","collide_ms","client_and_python_ms","engine_frames","prefetch_hits","vlm_bad_images","returncode")})
PY
