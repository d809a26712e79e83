#!/bin/bash
# DFT_RFFT on the tensor cores: operands generated in the kernel (BPV_DFT_TMA=0) against TMA-fed operand images (default)
tag=${1:-r2j}
o=gpurun_out/${tag}_dft_tc.txt; : > $o
timeout 600 python -m pytest tests/test_window_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py -m gpu -q -k "dft or DFT or golden or replay" 2>&1 | tail -3 | tee -a $o
for v in 0 1; do
  echo "== BPV_DFT_TMA=$v config-4 shape: 1024 streams, W=1200, INTERP_CUBIC + FILTER_BUTTER, DFT_RFFT" >> $o
  BPV_DFT_TMA=$v timeout 300 python tools/bench_window.py --S 1024 --T 1 --W 1200 --fps 120 --methods INTERP_CUBIC,FILTER_BUTTER --transform DFT_RFFT --irregular --windows last 2>&1 | tail -6 >> $o
  echo "== BPV_DFT_TMA=$v 8192 streams, W=300, INTERP_LINEAR, DFT_RFFT" >> $o
  BPV_DFT_TMA=$v timeout 300 python tools/bench_window.py --S 8192 --T 1 --W 300 --methods INTERP_LINEAR --transform DFT_RFFT --irregular --windows last 2>&1 | tail -6 >> $o
done
cat $o
for v in 0 1; do
  echo "== bench.py BPV_XC_FIRST=$v"
  BPV_XC_FIRST=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --no-other 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step',d['ms_per_step'],{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()},'frac',d['roofline']['frac'])"
done
