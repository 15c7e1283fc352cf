#!/bin/bash
# GPU round-2 run EE (1 GPU): cProfile of the image-generation seam; tests + default bench after the trip-width rule change.
mkdir -p gpurun_out
T=r02ee
timeout 600 python tools/profile_seams.py > gpurun_out/${T}_profile_seams.txt 2> gpurun_out/${T}_profile_seams.err
cat gpurun_out/${T}_profile_seams.txt | head -60
tail -c 300 gpurun_out/${T}_profile_seams.err
