#!/bin/bash
# One gpurun call that regenerates the round's 1-GPU evidence under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --semantics reference_compiled > gpurun_out/bench_ours_rc.json 2> gpurun_out/bench_ours_rc.err
python bench.py --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
# launch list of the SAME command as the bench line (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/b_ncu.log 2>&1
# the dominant kernel of the bench step, one launch, full set
ncu --set full --clock-control none --import-source on -k regex:nlm_tiled -s 3 -c 1 -f -o gpurun_out/prof_bench \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_bench.log 2>&1
# the two kernel families added in round 2: float64 instantiation, box-mean fast path
ncu --set full --clock-control none --import-source on -k regex:nlm_tiled -c 1 -f -o gpurun_out/prof_f64 \
    python tools/dev_generic.py > gpurun_out/ncu_f64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:boxmean -s 6 -c 2 -f -o gpurun_out/prof_boxmean \
    python bench.py --semantics reference_compiled --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_boxmean.log 2>&1
tail -2 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -8
python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json')); r=json.load(open('gpurun_out/bench_ref.json'))
print('ours value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'clocks', d['clocks'], 'parity', d['parity']['max_scaled_err'])
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], '| ref arm', r['value'], r['cpu_baseline']['cores'])"
