#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/api_profile.py 32000000 8.5 > gpurun_out/r2r_api_profile.txt 2>&1
head -90 gpurun_out/r2r_api_profile.txt | cut -c1-170
