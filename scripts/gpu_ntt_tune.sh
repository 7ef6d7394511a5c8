#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ntt_gpu.py -x -q -m gpu 2>&1 | tail -3
for t in 11 10 9; do
  echo "== PB200_NTT_TILE_LOG=$t"
  PB200_NTT_TILE_LOG=$t SIZES=16,18,20,22,24,26 NTT_ONLY=1 python scripts/sweep.py 2>&1 >/dev/null | grep fft_ms | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['log_n'], 'fft %.4f ms  ifft %.4f  cfft %.4f  cifft %.4f  imad_frac %.3f'%(r['fft_ms'],r['ifft_ms'],r['coset_fft_ms'],r['coset_ifft_ms'],r['imad_frac']))"
done
