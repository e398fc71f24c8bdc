set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --workload cfg5 --rows 64 --steps 2 --warmup 3 > gpurun_out/r2_bench_cfg5_8gpu.json 2> gpurun_out/r2_bench_cfg5_8gpu.err; tail -c 300 gpurun_out/r2_bench_cfg5_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg3_8gpu.json 2> gpurun_out/r2_bench_cfg3_8gpu.err; tail -c 300 gpurun_out/r2_bench_cfg3_8gpu.err
python bench.py --apply-njobs 8 --workload cfg3 --rows 1024 --steps 2 --warmup 1 > gpurun_out/r2_apply_njobs8.json 2> gpurun_out/r2_apply_njobs8.err; tail -c 300 gpurun_out/r2_apply_njobs8.err
python bench.py --apply-njobs 1 --workload cfg3 --rows 1024 --steps 2 --warmup 1 > gpurun_out/r2_apply_njobs1.json 2> gpurun_out/r2_apply_njobs1.err
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r2_topo.txt; free -g >> gpurun_out/r2_topo.txt
cat gpurun_out/r2_apply_njobs8.json gpurun_out/r2_apply_njobs1.json
