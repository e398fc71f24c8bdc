#!/bin/bash
# One gpurun call that regenerates the round's 1-GPU evidence under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nlm_tiled -s 3 -c 1 -o gpurun_out/prof_bench \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -3
python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json')); r=json.load(open('gpurun_out/bench_ref.json'))
print('ours value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'clocks', d['clocks'])
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], '| ref arm', r['value'], r['cpu_baseline']['cores'])"
