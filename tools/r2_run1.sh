set -x
python bench.py --workload cfg1 --steps 5 > gpurun_out/r2_b_cfg1.json 2> gpurun_out/r2_b_cfg1.err; tail -c 600 gpurun_out/r2_b_cfg1.err
python bench.py --workload cfg2 --steps 5 --no-cpu > gpurun_out/r2_b_cfg2.json 2> gpurun_out/r2_b_cfg2.err; tail -c 600 gpurun_out/r2_b_cfg2.err
python bench.py --workload cfg4 --rows 512 --steps 2 --warmup 1 > gpurun_out/r2_b_cfg4.json 2> gpurun_out/r2_b_cfg4.err; tail -c 600 gpurun_out/r2_b_cfg4.err
python bench.py --workload cfg5 --rows 64 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_b_cfg5.json 2> gpurun_out/r2_b_cfg5.err; tail -c 600 gpurun_out/r2_b_cfg5.err
