#!/bin/bash
mkdir -p gpurun_out
for f in 0 1; do
echo "== PB200_NTT_FUSED=$f"
PB200_NTT_FUSED=$f timeout 900 python -m pytest tests/test_ntt_gpu.py -x -q -m gpu 2>&1 | tail -1
PB200_NTT_FUSED=$f SIZES=16,20,22,24 NTT_ONLY=1 python scripts/sweep.py 2>&1 >/dev/null | grep fft_ms | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['log_n'], 'fft %.4f ms  ifft %.4f  cfft %.4f  cifft %.4f  imad_frac %.3f'%(r['fft_ms'],r['ifft_ms'],r['coset_fft_ms'],r['coset_ifft_ms'],r['imad_frac']))"
done
