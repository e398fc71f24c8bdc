#!/bin/bash
# One short gpurun call: GPU tests + smoke with the shipped binary, `ncu --set full` of the two f = 2 instantiations
# that BASELINE configs[3] / [4] run (cfg4: V = 4, r = (7,7,2); cfg5: V = 6, r = (10,10,3)) on development shapes,
# and fresh bench lines of the small workloads.  Outputs under gpurun_out/ (summaries are copied into profiles/).
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python tools/dev_bench.py 256 512 64 4 7 7 2 2 > gpurun_out/dev_cfg4.txt 2>&1
python tools/dev_bench.py 64 256 128 6 10 10 3 2 > gpurun_out/dev_cfg5.txt 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:nlm_tiled -s 1 -c 1 -f -o gpurun_out/prof_cfg4_f2 \
    python tools/dev_bench.py 256 512 64 4 7 7 2 2 --steps 1 > gpurun_out/ncu_cfg4.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:nlm_tiled -s 1 -c 1 -f -o gpurun_out/prof_cfg5_f2 \
    python tools/dev_bench.py 64 256 128 6 10 10 3 2 --steps 1 > gpurun_out/ncu_cfg5.log 2>&1
timeout 150 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 150 python bench.py --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
tail -3 gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/smoke.txt; cat gpurun_out/dev_cfg4.txt gpurun_out/dev_cfg5.txt | cut -c1-400
ls -la gpurun_out
