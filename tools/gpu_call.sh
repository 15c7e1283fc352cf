mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/r01k_pytest.log 2>&1
tail -4 gpurun_out/r01k_pytest.log
timeout 900 python bench.py > gpurun_out/r01k_bench.json 2> gpurun_out/r01k_bench.err; tail -2 gpurun_out/r01k_bench.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r01k_bench_reference.json 2>> gpurun_out/r01k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --launch-skip 39 -c 26 --csv --log-file gpurun_out/r01k_launches.csv python tools/frame_prof.py > gpurun_out/r01k_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 52 -c 13 -o gpurun_out/r01k_frame_full -f python tools/frame_prof.py > gpurun_out/r01k_ncu.log 2>&1; tail -1 gpurun_out/r01k_ncu.log
timeout 600 python tools/episode_bench.py --steps 300 --capsule > gpurun_out/r01k_episode.json 2> gpurun_out/r01k_episode.err; tail -1 gpurun_out/r01k_episode.err
timeout 600 python tools/episode_bench.py --steps 300 --vlm-client > gpurun_out/r01k_episode_client.json 2>> gpurun_out/r01k_episode.err
timeout 600 python tools/episode_bench.py --steps 300 --vlm-client --jpeg-handoff > gpurun_out/r01k_episode_client_handoff.json 2>> gpurun_out/r01k_episode.err
timeout 600 python tools/plan_bench.py > gpurun_out/r01k_plan_bench.json 2> gpurun_out/r01k_plan.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r01k_bench.json"))
print("value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "handoff", round(j["handoff_jpeg"]["value"],1), "roofline", j["roofline"]["kernel"], round(j["roofline"]["frac"],4))
for n in ("episode","episode_client","episode_client_handoff"):
    e=json.load(open(f"gpurun_out/r01k_{n}.json")); print(n, round(e["steps_per_s"],1), "render_ms", round(e["render_ms"],2), "client_ms", e.get("vlm_client_ms"))
print(open("gpurun_out/r01k_plan_bench.json").read()[:400])
PY
