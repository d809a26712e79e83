#!/bin/bash
# r4a: packed-float xcorr / LS kernels, F2 with global taps at 80 registers + dense-window fast path: parity, bench A/B, host profile
tag=r4a
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log | cut -c1-220
python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
BPV_FIR_TAPS_SMEM=1 python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2_taps_smem.json 2>> gpurun_out/${tag}_bench.err; echo "bench2 rc=$?"
python tools/profile_process.py 200 > gpurun_out/${tag}_profile_process.txt 2>&1; echo "prof rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_c2.json','gpurun_out/${tag}_bench_c2_taps_smem.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
    print({k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    print('roofline',d['roofline']['frac'],'by_time',d['roofline_by_time'].get('frac'))
    for k,v in d.get('other_shapes',{}).items(): print(k,{a:b for a,b in v.items() if a not in ('desc','cpu')})
    print('lat',d.get('latency_c1'))
PY
head -60 gpurun_out/${tag}_profile_process.txt
