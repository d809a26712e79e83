#!/bin/bash
# DFT on the tensor cores: samples converted by the CTA's producer warps (BPV_DFT_SPLIT=0) against ready-made A images (default)
tag=${1:-r3d}
o=gpurun_out/${tag}_dft_tc.txt; : > $o
timeout 600 python -m pytest tests/test_window_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py tests/test_properties_gpu.py -m gpu -q -k "dft or DFT or golden or replay or cubic" 2>&1 | tail -3 | tee -a $o
C4="--S 1024 --T 1 --W 1200 --fps 120 --methods INTERP_CUBIC,FILTER_BUTTER --transform DFT_RFFT --irregular --windows last"
W3="--S 8192 --T 1 --W 300 --methods INTERP_LINEAR --transform DFT_RFFT --irregular --windows last"
for v in 0 1; do
  echo "== BPV_DFT_SPLIT=$v config-4 shape" >> $o
  BPV_DFT_SPLIT=$v timeout 300 python tools/bench_window.py $C4 2>&1 | tail -6 >> $o
  echo "== BPV_DFT_SPLIT=$v 8192 streams, W=300" >> $o
  BPV_DFT_SPLIT=$v timeout 300 python tools/bench_window.py $W3 2>&1 | tail -6 >> $o
  BPV_DFT_SPLIT=$v timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:dft_ -c 72 --csv --log-file gpurun_out/${tag}_dft_launches_c4_split$v.csv python tools/bench_window.py $C4 --iters 3 > /dev/null 2>&1
  BPV_DFT_SPLIT=$v timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:dft_ -c 72 --csv --log-file gpurun_out/${tag}_dft_launches_w300_split$v.csv python tools/bench_window.py $W3 --iters 3 > /dev/null 2>&1
done
cat $o | grep "passed\|failed\|rror\|window_spectrum\|=="
python - <<PY
import csv,collections
for shp in ('c4','w300'):
  for v in ('0','1'):
    rows=[l for l in open('gpurun_out/${tag}_dft_launches_%s_split%s.csv'%(shp,v)) if not l.startswith('==')]
    agg=collections.OrderedDict()
    for r in csv.DictReader(rows):
        k=(r['Kernel Name'].split('(')[0][-40:], r['Metric Name'][:30])
        try: agg.setdefault(k,[]).append(float(r['Metric Value'].replace(',','')))
        except ValueError: pass
    for k,vals in agg.items(): print(shp,'SPLIT=%s'%v, k, 'n=%d median=%.2f min=%.2f'%(len(vals), sorted(vals)[len(vals)//2], min(vals)))
PY
