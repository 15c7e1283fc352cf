#!/bin/bash
# GPU round-2 run H (1 GPU): tests after the seam work (RGBA shim, JPEG from RGBA, twin), facade in the trajectory frame, image generation.
mkdir -p gpurun_out
T=r02h
timeout 1200 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 600 python tools/episode_bench.py --steps 500 --capsule > gpurun_out/${T}_facade_1080p.json 2> gpurun_out/${T}_facade_1080p.err
timeout 600 python tools/episode_bench.py --steps 500 --capsule --width 640 --height 480 > gpurun_out/${T}_facade_480p.json 2> gpurun_out/${T}_facade_480p.err
timeout 600 python tools/episode_bench.py --steps 300 --capsule --width 640 --height 480 --stage-timing > gpurun_out/${T}_facade_480p_stages.json 2> gpurun_out/${T}_facade_480p_stages.err
timeout 600 python tools/imagegen_bench.py > gpurun_out/${T}_imagegen.json 2> gpurun_out/${T}_imagegen.err
for f in gpurun_out/${T}_facade*.json gpurun_out/${T}_imagegen.json; do echo $f; tail -1 $f | cut -c1-1600; done
tail -c 600 gpurun_out/${T}_*.err
