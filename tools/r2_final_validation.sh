#!/bin/bash
# Last gpurun call of round 2: the GPU test suite, smoke() and the default bench line with the final binary
# (double-duty halo warps for f = 2 / float64, row-blocked staging kernel, kernels writing native outputs in place).
mkdir -p gpurun_out
( timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/pytest_gpu_final.txt 2>&1
tail -3 gpurun_out/pytest_gpu_final.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; tail -2 gpurun_out/smoke_final.txt
timeout 60 python tools/dev_bench.py 420 4096 32 4 5 5 2 1 > gpurun_out/dev_cfg3_final.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/dev_cfg3_final.txt'):
    if l.startswith('BENCH'):
        d = json.loads(l[6:]); print('dev cfg3: stage %.3f ms  run %.3f ms  unstage %.3f ms' % (d['ms_stage'], d['ms_run'], d['ms_unstage']))
PY
timeout 200 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_cfg3_final.json 2> gpurun_out/bench_cfg3_final.err
tail -c 300 gpurun_out/bench_cfg3_final.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cfg3_final.json'))
print('cfg3 value', d['value'], 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', d['parity']['max_scaled_err'], 'launches', d['gpu_launches'])
PY
