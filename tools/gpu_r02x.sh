#!/bin/bash
# GPU round-2 run X (1 GPU): composite walk restructured (scan -> id ring -> full 32-entry stages) vs the chunk-per-stage loop of run W;
# two-entries-per-trip variant; parity of all; per-warp walk statistics of the new loops at 480p.
mkdir -p gpurun_out
T=r02x
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -3 > gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_pytest.log
for v in pair scan2; do
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 900 python -m pytest tests/test_render_parity_gpu.py -m gpu -q --no-header -x 2>&1 | tail -2 > gpurun_out/${T}_pytest_$v.log
  tail -1 gpurun_out/${T}_pytest_$v.log
done
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "640 480" "1024 768" "1920 1080"; do
  set -- $res
  timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_new_$2.json 2> gpurun_out/${T}_bench_new_$2.err
  for v in old pair scan2; do
    SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_${v}_$2.json 2> gpurun_out/${T}_bench_${v}_$2.err
  done
done
for v in prof pairprof; do
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_$v.so timeout 300 python tools/composite_walk_prof.py --width 640 --height 480 > gpurun_out/${T}_walk_${v}_480.json 2> gpurun_out/${T}_walk_${v}_480.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02x_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3), "composite", round(d["roofline"]["frame"]["stage_ms_serial"]["composite"],4), "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4))
    except Exception as e:
        print(f, "ERR", e)
for v in ("prof","pairprof"):
    try:
        d=json.load(open(f"gpurun_out/r02x_walk_{v}_480.json"))
        for p in d["poses"]:
            t=p["top5"][0]
            print(v, "pose",p["pose"],"span",round(p["kernel_span_us"],1),"end p50/90/99",[round(x,1) for x in p["end_us_p50_p90_p99"]],"visited",p["visited_sum"],"hits",p["hits_sum"],"scanned",p["scanned_sum"],"cyc_sum",p["cycles_sum"], "top:", {k:(round(x,1) if isinstance(x,float) else x) for k,x in t.items()})
    except Exception as e:
        print(v, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench_new_480.err
