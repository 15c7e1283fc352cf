#!/bin/bash
# GPU round-2 run Y (1 GPU): k_composite with E entries per trip (E = 2 large frames, 4 small frames) - parity of both, bench at three
# resolutions with each E forced, per-warp walk statistics at 480p; default pipeline depth 5.
mkdir -p gpurun_out
T=r02y
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "640 480" "1024 768" "1920 1080"; do
  set -- $res
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_old.so SAGE_PIPELINE_DEPTH=5 timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_old_$2.json 2> gpurun_out/${T}_bench_old_$2.err
  SAGE_COMPOSITE_WIDE_MAX_CTAS=0 timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_e2_$2.json 2> gpurun_out/${T}_bench_e2_$2.err
  SAGE_COMPOSITE_WIDE_MAX_CTAS=100000000 timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_e4_$2.json 2> gpurun_out/${T}_bench_e4_$2.err
done
SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_prof.so timeout 300 python tools/composite_walk_prof.py --width 640 --height 480 > gpurun_out/${T}_walk_e4_480.json 2> gpurun_out/${T}_walk_e4_480.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02y_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "composite", round(d["roofline"]["frame"]["stage_ms_serial"]["composite"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
try:
    d=json.load(open(f"gpurun_out/r02y_walk_e4_480.json"))
    for p in d["poses"]:
        t=p["top5"][0]
        print("e4 pose",p["pose"],"span",round(p["kernel_span_us"],1),"end p50/90/99",[round(x,1) for x in p["end_us_p50_p90_p99"]],"visited",p["visited_sum"],"hits",p["hits_sum"],"cyc_sum",p["cycles_sum"], "top:", {k:(round(x,1) if isinstance(x,float) else x) for k,x in t.items()})
except Exception as e:
    print("ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench_e4_480.err
