#!/bin/bash
# r4b: design probe on a high-priority stream, hole-free fast paths in Welch / xcorr, drop-in process() host path: parity, bench, A/B of the priority
tag=r4b
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log | cut -c1-220
python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
BPV_DESIGN_PRIO=0 python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2_prio0.json 2>> gpurun_out/${tag}_bench.err; echo "bench2 rc=$?"
python tools/profile_process.py 200 > gpurun_out/${tag}_profile_process.txt 2>&1; echo "prof rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_c2.json','gpurun_out/${tag}_bench_c2_prio0.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
    print({k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    print('roofline',d['roofline']['frac'],'by_time',d['roofline_by_time'].get('frac'))
    for k,v in d.get('other_shapes',{}).items(): print(k,v['ms_per_step'])
    print('lat',d.get('latency_c1'))
PY
head -45 gpurun_out/${tag}_profile_process.txt
