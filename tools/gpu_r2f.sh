#!/bin/bash
mkdir -p gpurun_out
python tools/c3_variants.py 2>&1 | tail -12 | tee gpurun_out/r2f_c3_variants.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2f_pytest.log
