#!/bin/bash
# batched-scan tests + tensor-core kernel timing (mask off / on, several thresholds)
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -k "batched" 2>&1 | tail -5 | tee gpurun_out/c5_pytest.log
python tools/c5_perf.py 125e6 2>&1 | tail -8 | tee gpurun_out/c5_perf.txt
