python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compiled" 2>&1 | tail -3
python bench.py --workload cfg3 --semantics reference_compiled --steps 5 --no-e2e --no-cpu > gpurun_out/r2_b_cfg3_rc.json 2> gpurun_out/r2_b_cfg3_rc.err; tail -c 400 gpurun_out/r2_b_cfg3_rc.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:boxmean -s 4 -c 4 --csv --log-file gpurun_out/r2_rc_launches.csv python bench.py --workload cfg3 --semantics reference_compiled --steps 1 --no-e2e --no-cpu --no-parity > /dev/null 2>&1
tail -14 gpurun_out/r2_rc_launches.csv
