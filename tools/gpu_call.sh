mkdir -p gpurun_out
python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/r01b_pytest.log 2>&1
tail -40 gpurun_out/r01b_pytest.log
timeout 600 python bench.py --steps 500 --warmup 50 > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err; tail -3 gpurun_out/r01b_bench.err
timeout 300 python tools/jpeg_prof.py > gpurun_out/r01b_jpeg_prof.log 2>&1; cat gpurun_out/r01b_jpeg_prof.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_jpeg --launch-skip 12 -c 6 -o gpurun_out/r01b_jpeg_full -f python tools/jpeg_prof.py > gpurun_out/r01b_ncu.log 2>&1; tail -3 gpurun_out/r01b_ncu.log
head -c 6000 gpurun_out/r01b_bench.json
