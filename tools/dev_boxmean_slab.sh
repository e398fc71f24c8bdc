for S in 0 22 38 54 118 246; do
  NDNLM_BOXMEAN_SLAB=$S python bench.py --semantics reference_compiled --steps 5 --no-e2e --no-cpu > gpurun_out/r2_box_$S.json 2> gpurun_out/r2_box_$S.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_box_$S.json'))
print('slab', $S, 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'step ms', round(d['ms_per_step'],2), 'parity', d['parity']['max_scaled_err'], 'launches', d['gpu_launches'])"
done
NDNLM_BOXMEAN_SLAB=54 python -m pytest tests/test_gpu_parity.py -q -m gpu -k compiled 2>&1 | tail -2
