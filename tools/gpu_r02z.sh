#!/bin/bash
# GPU round-2 run Z (1 GPU): radix-pass histograms accumulated by the producers (no histogram kernels) vs the build before it;
# frozen-spec digests on the GPU; A* slots / memory budget sweep; turnover of one rank's block of an 8-way split.
mkdir -p gpurun_out
T=r02z
timeout 1500 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; tail -1 gpurun_out/${T}_pytest.log
B="python bench.py --steps 200 --warmup 30 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
for res in "640 480" "1920 1080"; do
  set -- $res
  SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_nofuse.so timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_nofuse_$2.json 2> gpurun_out/${T}_bench_nofuse_$2.err
  timeout 300 $B --width $1 --height $2 > gpurun_out/${T}_bench_fused_$2.json 2> gpurun_out/${T}_bench_fused_$2.err
done
timeout 300 python tools/strip_rate.py --split 8 --ctx 6 > gpurun_out/${T}_strip8_ctx6.json 2> gpurun_out/${T}_strip8_ctx6.err
SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_nofuse.so timeout 300 python tools/strip_rate.py --split 8 --ctx 6 > gpurun_out/${T}_strip8_ctx6_nofuse.json 2> gpurun_out/${T}_strip8_ctx6_nofuse.err
cat gpurun_out/${T}_strip8_ctx6_nofuse.json gpurun_out/${T}_strip8_ctx6.json
timeout 200 python tools/plan_bench.py --slots 4096 --budget-gb 4 > gpurun_out/${T}_plan_4k.json 2> gpurun_out/${T}_plan_4k.err
timeout 200 python tools/plan_bench.py --slots 16384 --budget-gb 20 > gpurun_out/${T}_plan_16k.json 2> gpurun_out/${T}_plan_16k.err
timeout 200 python tools/plan_bench.py --slots 32768 --budget-gb 40 --queries 32768 > gpurun_out/${T}_plan_32k.json 2> gpurun_out/${T}_plan_32k.err
cat gpurun_out/${T}_plan_4k.json gpurun_out/${T}_plan_16k.json gpurun_out/${T}_plan_32k.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02z_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"lat",round(d["frame_latency_ms"],3))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 300 gpurun_out/${T}_bench_fused_1080.err; tail -c 300 gpurun_out/${T}_plan_32k.err
