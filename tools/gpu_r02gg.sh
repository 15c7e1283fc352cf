#!/bin/bash
mkdir -p gpurun_out
echo "== PREFILTER=0"; SAGE_COMPOSITE_PREFILTER=0 timeout 600 python tools/diag_needles.py 2>&1 | tail -12
echo "== PREFILTER=1"; SAGE_COMPOSITE_PREFILTER=1 timeout 600 python tools/diag_needles.py 2>&1 | tail -12
