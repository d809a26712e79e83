#!/bin/bash
# DFT_RFFT on the tensor cores: operands generated in the kernel (BPV_DFT_TMA=0) against the TMA-fed, warp-specialised kernel (default)
tag=${1:-r2m}
o=gpurun_out/${tag}_dft_tc.txt; : > $o
timeout 600 python -m pytest tests/test_window_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py tests/test_properties_gpu.py -m gpu -q -k "dft or DFT or golden or replay or cubic" 2>&1 | tail -3 | tee -a $o
C4="--S 1024 --T 1 --W 1200 --fps 120 --methods INTERP_CUBIC,FILTER_BUTTER --transform DFT_RFFT --irregular --windows last"
W3="--S 8192 --T 1 --W 300 --methods INTERP_LINEAR --transform DFT_RFFT --irregular --windows last"
for v in 0 1; do
  echo "== BPV_DFT_TMA=$v config-4 shape: 1024 streams, W=1200, INTERP_CUBIC + FILTER_BUTTER, DFT_RFFT" >> $o
  BPV_DFT_TMA=$v timeout 300 python tools/bench_window.py $C4 2>&1 | tail -6 >> $o
  echo "== BPV_DFT_TMA=$v 8192 streams, W=300, INTERP_LINEAR, DFT_RFFT" >> $o
  BPV_DFT_TMA=$v timeout 300 python tools/bench_window.py $W3 2>&1 | tail -6 >> $o
  BPV_DFT_TMA=$v timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
      --clock-control none -k regex:dft_ -c 60 --csv --log-file gpurun_out/${tag}_dft_launches_tma$v.csv python tools/bench_window.py $C4 --iters 3 > /dev/null 2>&1
done
cat $o
python - <<PY
import csv,collections
for v in (0,1):
    rows=[l for l in open('gpurun_out/${tag}_dft_launches_tma%d.csv'%v) if not l.startswith('==')]
    agg=collections.OrderedDict()
    for r in csv.DictReader(rows):
        k=(r['Kernel Name'].split('(')[0][-40:], r['Metric Name'])
        try: agg.setdefault(k,[]).append(float(r['Metric Value'].replace(',','')))
        except ValueError: pass
    for k,vals in agg.items(): print('TMA=%d'%v, k, 'n=%d median=%.2f min=%.2f'%(len(vals), sorted(vals)[len(vals)//2], min(vals)))
PY
echo "== bench.py (timed region with F1 events only)"
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-other 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step',d['ms_per_step'],{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()},'frac',d['roofline']['frac'])"
