#!/bin/bash
# GPU round-2 run K (4 GPUs): two-GPU frame-equality test with the flag barrier, scaling at N=4 (flags vs NCCL barrier, 3 vs 6 contexts), N=2.
mkdir -p gpurun_out
T=r02k
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --no-header -x 2>&1 | tail -8 > gpurun_out/${T}_pytest_mg.log
tail -4 gpurun_out/${T}_pytest_mg.log
B="bench.py --steps 100 --warmup 20 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff"
run() { # name gpus
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 $B --gpus $2 > gpurun_out/${T}_$1.json 2> gpurun_out/${T}_$1.err
}
run n4_flags 4
SAGE_PEER_BARRIER=nccl run n4_nccl 4
SAGE_SPLAT_LIB=sage-3d_official_b200/libsage_splat_ctx6.so run n4_ctx6 4
run n2_flags 2
timeout 600 python $B > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02k_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1), "equal", d.get("multigpu_frame_equal"), d["e2e"].get("frame_equal"))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()}, "sum", round(d["roofline"]["frame"]["serial_frame_ms"],4), "lat", round(d["frame_latency_ms"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/${T}_n4_flags.err
