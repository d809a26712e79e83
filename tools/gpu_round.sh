#!/bin/bash
# One GPU session: parity tests, the bench line, the ncu launch list of the bench command, one --set full capture of a step.
# Usage: tools/gpu_round.sh <tag> [full]
tag=${1:-r2a}; full=${2:-}
mkdir -p gpurun_out
KF='regex:roi_|ring_push|firls|butter|window_preprocess|spectrum_dense|welch|xcorr|ls_coarse|ls_peak|running_mean|calc_rois|pack_records|dft_|design_probe|miss_'
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log | cut -c1-220
python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 --ref-budget-s 20 > gpurun_out/${tag}_bench_c2_reference_arm.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-other > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
if [ -n "$full" ]; then
  # ring prefill = 5 step_signals x 5 kernels; each step = 9 kernels -> skip past the prefill and two warm-up steps, capture
  # two whole steps
  ncu --set full --clock-control none --import-source on -k "$KF" --launch-skip 43 --launch-count 18 -f -o gpurun_out/${tag}_c2 \
      python bench.py --steps 1 --warmup 3 --no-cpu --no-other > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_c2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
print('roofline',d['roofline']['frac'],'by_time',d['roofline_by_time'].get('frac'),'fma',d.get('fma_peaks_tflops'))
print('cpu',d.get('cpu_baseline',{}).get('value'),d.get('cpu_baseline',{}).get('kind'))
for k,v in d.get('other_shapes',{}).items(): print(k,{a:b for a,b in v.items() if a not in ('desc','cpu')})
print('lat',d.get('latency_c1'))
PY
