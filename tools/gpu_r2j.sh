#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_fullsize_gpu.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2j_pytest_fullsize.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2j_smoke.log
python tools/api_profile.py 32e6 2>&1 | head -45 | tee gpurun_out/r2j_api_profile.txt
