#!/bin/bash
# GPU round-2 run F (2 GPUs): the two-GPU frame-equality test, bench at N=2 and N=1.
mkdir -p gpurun_out
T=r02f
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --no-header -x 2>&1 | tail -30 > gpurun_out/${T}_pytest_mg.log
tail -6 gpurun_out/${T}_pytest_mg.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
tail -c 1500 gpurun_out/${T}_bench_2gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch-scenes 0 --no-collide --no-handoff > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value",round(d["value"],1),"e2e",d["e2e"],"clocks",d["clocks"])
        for k in ("multigpu_frame_equal","multigpu","batch8","timed_blocks"):
            print("  ",k,d.get(k))
        print("  serial",{k:round(v,4) for k,v in d["roofline"]["frame"]["stage_ms_serial"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
