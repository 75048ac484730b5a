#!/bin/bash
# quick check of one test file / expression:  bash tools/gpu_quick.sh <pytest args>
mkdir -p gpurun_out
python -m pytest "$@" -x -q 2>&1 | tail -25 | tee gpurun_out/quick_pytest.log
