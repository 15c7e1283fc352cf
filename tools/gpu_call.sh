mkdir -p gpurun_out
timeout 600 python bench.py --steps 500 --warmup 50 --no-cpu-baseline > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err; tail -2 gpurun_out/r01j_bench.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r01j_bench.json"))
print("value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "V", j["config"]["mean_visible"], "D", j["config"]["mean_instances"]); print(j["roofline"]["frame"]["stage_ms_serial"])
PY
