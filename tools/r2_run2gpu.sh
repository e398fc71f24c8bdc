set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two or nccl" 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload cfg4 --global-rows 1024 --steps 2 --warmup 1 --no-e2e > gpurun_out/r2_cfg4_dev_2gpu.json 2> gpurun_out/r2_cfg4_dev_2gpu.err; tail -c 500 gpurun_out/r2_cfg4_dev_2gpu.err
python bench.py --apply-njobs 2 --workload cfg3 --rows 1024 --steps 2 --warmup 1 > gpurun_out/r2_apply_njobs2.json 2> gpurun_out/r2_apply_njobs2.err; tail -c 500 gpurun_out/r2_apply_njobs2.err
python bench.py --apply-njobs 1 --workload cfg3 --rows 1024 --steps 2 --warmup 1 > gpurun_out/r2_apply_njobs1.json 2> gpurun_out/r2_apply_njobs1.err; tail -c 500 gpurun_out/r2_apply_njobs1.err
cat gpurun_out/r2_apply_njobs2.json gpurun_out/r2_apply_njobs1.json
