#!/bin/bash
# TMA-fed DFT kernel: 128 bins per CTA (BPV_DFT_BC=128) against the grid-filling bin chunk width (default)
tag=${1:-r2s}
o=gpurun_out/${tag}_dft_tc.txt; : > $o
timeout 600 python -m pytest tests/test_window_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py tests/test_properties_gpu.py -m gpu -q -k "dft or DFT or golden or replay or cubic" 2>&1 | tail -3 | tee -a $o
C4="--S 1024 --T 1 --W 1200 --fps 120 --methods INTERP_CUBIC,FILTER_BUTTER --transform DFT_RFFT --irregular --windows last"
W3="--S 8192 --T 1 --W 300 --methods INTERP_LINEAR --transform DFT_RFFT --irregular --windows last"
for v in 128 auto; do
  echo "== BPV_DFT_BC=$v config-4 shape: 1024 streams, W=1200, INTERP_CUBIC + FILTER_BUTTER, DFT_RFFT" >> $o
  BPV_DFT_BC=$v timeout 300 python tools/bench_window.py $C4 2>&1 | tail -6 >> $o
  echo "== BPV_DFT_BC=$v 8192 streams, W=300, INTERP_LINEAR, DFT_RFFT" >> $o
  BPV_DFT_BC=$v timeout 300 python tools/bench_window.py $W3 2>&1 | tail -6 >> $o
  BPV_DFT_BC=$v timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
      --clock-control none -k regex:dft_ -c 60 --csv --log-file gpurun_out/${tag}_dft_launches_c4_bc$v.csv python tools/bench_window.py $C4 --iters 3 > /dev/null 2>&1
  BPV_DFT_BC=$v timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:dft_ -c 60 --csv --log-file gpurun_out/${tag}_dft_launches_w300_bc$v.csv python tools/bench_window.py $W3 --iters 3 > /dev/null 2>&1
done
cat $o
python - <<PY
import csv,collections
for shp in ('c4','w300'):
  for v in ('128','auto'):
    rows=[l for l in open('gpurun_out/${tag}_dft_launches_%s_bc%s.csv'%(shp,v)) if not l.startswith('==')]
    agg=collections.OrderedDict()
    for r in csv.DictReader(rows):
        k=(r['Kernel Name'].split('(')[0][-40:], r['Metric Name'])
        try: agg.setdefault(k,[]).append(float(r['Metric Value'].replace(',','')))
        except ValueError: pass
    for k,vals in agg.items(): print(shp,'BC=%s'%v, k, 'n=%d median=%.2f min=%.2f'%(len(vals), sorted(vals)[len(vals)//2], min(vals)))
PY
