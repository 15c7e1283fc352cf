mkdir -p gpurun_out
N=${SAGE_N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 50 > gpurun_out/r01k_bench_${N}gpu.json 2> gpurun_out/r01k_bench_${N}gpu.err
tail -3 gpurun_out/r01k_bench_${N}gpu.err; head -c 700 gpurun_out/r01k_bench_${N}gpu.json
