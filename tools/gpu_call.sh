mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 600 -x > gpurun_out/r01d_pytest.log 2>&1
tail -5 gpurun_out/r01d_pytest.log
SAGE_NO_FORK=1 timeout 600 python bench.py --steps 500 --warmup 50 --no-cpu-baseline > gpurun_out/r01d_bench_nofork.json 2> gpurun_out/r01d_bench_nofork.err; tail -2 gpurun_out/r01d_bench_nofork.err
timeout 600 python bench.py --steps 500 --warmup 50 --no-cpu-baseline > gpurun_out/r01d_bench_fork.json 2> gpurun_out/r01d_bench_fork.err; tail -2 gpurun_out/r01d_bench_fork.err
python - <<'PY'
import json
for n in ("nofork","fork"):
    try:
        j=json.load(open(f"gpurun_out/r01d_bench_{n}.json"))
        print(n, "value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "handoff", round(j["handoff_jpeg"]["value"],1), "enc_ms", round(j["handoff_jpeg"]["jpeg_encode_ms_device"],4), j["roofline"]["frame"]["stage_ms"])
    except Exception as e: print(n, "ERR", e)
PY
