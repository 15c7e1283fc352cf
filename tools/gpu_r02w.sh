#!/bin/bash
# GPU round-2 run W (1 GPU): per-warp walk statistics of k_composite (profile variant) at 480p / 768p / 1080p.
mkdir -p gpurun_out
T=r02w
export SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_prof.so
for res in "640 480" "1024 768" "1920 1080"; do
  set -- $res
  timeout 300 python tools/composite_walk_prof.py --width $1 --height $2 > gpurun_out/${T}_walk_$2.json 2> gpurun_out/${T}_walk_$2.err
  tail -c 200 gpurun_out/${T}_walk_$2.err
done
python - <<'PY'
import json
for r in (480,768,1080):
    d=json.load(open(f"gpurun_out/r02w_walk_{r}.json"))
    for p in d["poses"]:
        t=p["top5"][0]
        print(r, "pose",p["pose"],"span",round(p["kernel_span_us"],1),"last_start",round(p["last_start_us"],1),"end p50/90/99",[round(x,1) for x in p["end_us_p50_p90_p99"]],"visited",p["visited_sum"],"hits",p["hits_sum"],"scanned",p["scanned_sum"],"cyc_sum",p["cycles_sum"])
        print("    top:",t)
PY
