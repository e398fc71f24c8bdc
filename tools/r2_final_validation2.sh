#!/bin/bash
# The GPU test suite with the final binary, all tests (no -x).
mkdir -p gpurun_out
( timeout 150 python -m pytest tests -q -m gpu 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_final.txt 2>&1
tail -4 gpurun_out/pytest_gpu_final.txt
