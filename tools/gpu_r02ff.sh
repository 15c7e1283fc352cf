#!/bin/bash
# GPU round-2 run FF (1 GPU): k_composite with the per-stage box test of the entries against the live pixels (PRE) on / off:
# parity (all tests with it on), bench at three resolutions, facade at 640x480, per-warp walk statistics.
mkdir -p gpurun_out
T=r02ff
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -8 > gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "640 480" "1024 768" "1920 1080"; do
  set -- $res
  for pre in 0 1; do
    SAGE_COMPOSITE_PREFILTER=$pre timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_pre${pre}_$2.json 2>> gpurun_out/${T}_bench.err
  done
done
SAGE_COMPOSITE_PREFILTER=1 SAGE_COMPOSITE_WIDE_MAX_CTAS=0 timeout 300 $B --width 640 --height 480 > gpurun_out/${T}_bench_pre1_e2_480.json 2>> gpurun_out/${T}_bench.err
E="python tools/episode_bench.py --steps 500 --capsule --width 640 --height 480"
SAGE_COMPOSITE_PREFILTER=0 timeout 300 $E > gpurun_out/${T}_facade_480p_pre0.json 2>> gpurun_out/${T}_bench.err
SAGE_COMPOSITE_PREFILTER=1 timeout 300 $E > gpurun_out/${T}_facade_480p_pre1.json 2>> gpurun_out/${T}_bench.err
SAGE_COMPOSITE_PREFILTER=1 timeout 300 python tools/strip_rate.py --split 8 --ctx 6 > gpurun_out/${T}_strip8_pre1.json 2>> gpurun_out/${T}_bench.err
SAGE_COMPOSITE_PREFILTER=0 timeout 300 python tools/strip_rate.py --split 8 --ctx 6 > gpurun_out/${T}_strip8_pre0.json 2>> gpurun_out/${T}_bench.err
SAGE_COMPOSITE_PREFILTER=1 SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_prof.so timeout 300 python tools/composite_walk_prof.py --width 640 --height 480 > gpurun_out/${T}_walk_pre1_480.json 2>> gpurun_out/${T}_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ff_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "composite", round(d["roofline"]["frame"]["stage_ms_serial"]["composite"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r02ff_facade_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ("steps_per_s","ms_per_step","render_ms","collide_ms")})
for f in sorted(glob.glob("gpurun_out/r02ff_strip*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], {k:(round(v["device_fps"])) for k,v in d.items() if isinstance(v,dict)})
try:
    d=json.load(open("gpurun_out/r02ff_walk_pre1_480.json"))
    for p in d["poses"]:
        t=p["top5"][0]
        print("pre1 pose",p["pose"],"span",round(p["kernel_span_us"],1),"end p50/90/99",[round(x,1) for x in p["end_us_p50_p90_p99"]],"visited",p["visited_sum"],"hit-trips",p["hits_sum"],"cyc_sum",p["cycles_sum"], "top:", {k:(round(x,1) if isinstance(x,float) else x) for k,x in t.items()})
except Exception as e:
    print("ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench.err
