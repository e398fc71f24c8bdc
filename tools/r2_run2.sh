set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compiled or nan_footprint or golden" 2>&1 | tail -15
python bench.py --workload cfg3 --semantics reference_compiled --steps 5 --no-e2e > gpurun_out/r2_b_cfg3_rc.json 2> gpurun_out/r2_b_cfg3_rc.err; tail -c 400 gpurun_out/r2_b_cfg3_rc.err
python bench.py --workload cfg1 --steps 5 --no-cpu > gpurun_out/r2_b_cfg1.json 2> gpurun_out/r2_b_cfg1.err; tail -c 400 gpurun_out/r2_b_cfg1.err
