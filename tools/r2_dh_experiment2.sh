#!/bin/bash
# Second double-duty-halo-warp call: f = 2 variants (12 warps x L = 6 against 16 warps x L = 4), the float64
# instantiations, the GPU test suite with the shipped default (DH for f_W >= 2) and with NDNLM_DH=1, the cfg4 bench
# line (1 GPU, stated partial sweep) and an ncu capture of the DH kernel.  Every step under `timeout`.
mkdir -p gpurun_out
{
echo "== cfg4 parameters (f = 2): auto (shipped default), 44 = L6 x 12 warps (dh), 45 = L4 x 16 warps (dh); NDNLM_DH=0 = round-2 kernel"
timeout 120 python tools/dev_multi.py auto,44,45 --shape 240,512,64,4 --r 7,7,2 --f 2 --pshape 24,40,10,4 --steps 3 2>&1 | tail -3
NDNLM_DH=0 timeout 120 python tools/dev_multi.py auto --shape 240,512,64,4 --r 7,7,2 --f 2 --pshape 24,40,10,4 --steps 3 2>&1 | tail -1
echo "== cfg4 parameters with n_eff = 50: shipped default (dh) against NDNLM_DH=0"
timeout 120 python tools/dev_multi.py auto --shape 240,512,64,4 --r 7,7,2 --f 2 --pshape 24,40,10,4 --steps 2 --neff 50 2>&1 | tail -1
NDNLM_DH=0 timeout 120 python tools/dev_multi.py auto --shape 240,512,64,4 --r 7,7,2 --f 2 --pshape 24,40,10,4 --steps 2 --neff 50 2>&1 | tail -1
echo "== float64"
timeout 120 python tools/dev_dh64.py 2>&1 | tail -6
} > gpurun_out/dh_experiment2.txt 2>&1
cat gpurun_out/dh_experiment2.txt
( timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/pytest_gpu_final.txt 2>&1
( NDNLM_DH=1 timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/pytest_gpu_dh1.txt 2>&1
tail -2 gpurun_out/pytest_gpu_final.txt gpurun_out/pytest_gpu_dh1.txt
timeout 200 python bench.py --workload cfg4 --rows 500 --steps 2 --warmup 3 > gpurun_out/bench_cfg4_1gpu_partial.json 2> gpurun_out/bench_cfg4_1gpu_partial.err
tail -c 400 gpurun_out/bench_cfg4_1gpu_partial.err; head -c 300 gpurun_out/bench_cfg4_1gpu_partial.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:nlm_tiled -s 1 -c 1 -f -o gpurun_out/prof_cfg4_f2_dh \
    python tools/dev_bench.py 240 512 64 4 7 7 2 2 --steps 1 > gpurun_out/ncu_cfg4_dh.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
