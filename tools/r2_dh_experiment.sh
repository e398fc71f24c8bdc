#!/bin/bash
# Double-duty halo warps (instances_g6.inc) against the shipped instantiations: parity vs the oracle on a small cube,
# kernel timing on a development cube whose row count is a multiple of both tile heights, then the GPU test suite
# with the DH instantiations preferred.  Every step under `timeout` (a protocol bug would be a deadlock).
mkdir -p gpurun_out
{
for dh in 0 1; do
  echo "== NDNLM_DH=$dh cfg3 parameters"
  NDNLM_DH=$dh timeout 120 python tools/dev_multi.py auto --shape 420,4096,32,4 --steps 3 2>&1 | tail -2
done
for dh in 0 1; do
  echo "== NDNLM_DH=$dh cfg4 parameters (f = 2)"
  NDNLM_DH=$dh timeout 120 python tools/dev_multi.py auto --shape 240,512,64,4 --r 7,7,2 --f 2 --pshape 24,40,10,4 --steps 3 2>&1 | tail -2
done
echo "== NDNLM_DH=1 cfg3 parameters, n_eff = 50"
NDNLM_DH=1 timeout 120 python tools/dev_multi.py auto --shape 420,4096,32,4 --steps 2 --neff 50 2>&1 | tail -2
} > gpurun_out/dh_experiment.txt 2>&1
cat gpurun_out/dh_experiment.txt
( NDNLM_DH=1 timeout 240 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) > gpurun_out/pytest_gpu_dh1.txt 2>&1
cat gpurun_out/pytest_gpu_dh1.txt
