# usage: r2_run_cfg4.sh NGPU  -- full-size BASELINE configs[3] (16384 x 16384 x 64, strong scaling), y-sharded over NGPU ranks
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload cfg4 --steps 2 --warmup 3 > gpurun_out/r2_bench_cfg4_${N}gpu.json 2> gpurun_out/r2_bench_cfg4_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_cfg4_${N}gpu.err; head -c 700 gpurun_out/r2_bench_cfg4_${N}gpu.json
