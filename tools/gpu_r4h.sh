#!/bin/bash
# r4h: F2 with several consecutive signals per warp (BPV_F2_SPW): parity with 3 per warp forced everywhere, bench A/B (auto = 2 vs 1)
tag=r4h
mkdir -p gpurun_out
BPV_F2_SPW=3 python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest_spw3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_spw3.log
tail -5 gpurun_out/${tag}_pytest_spw3.log | cut -c1-220
python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
BPV_F2_SPW=1 python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2_spw1.json 2>> gpurun_out/${tag}_bench.err; echo "bench spw1 rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_c2.json','gpurun_out/${tag}_bench_c2_spw1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
        print({k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'F1',d['roofline']['frac'],'F2',d['roofline_by_time'].get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
