#!/bin/bash
# r4g: Welch + xcorr as ONE grid of interleaved CTAs (overlap bit 4): parity with either schedule as the default, bench A/B
tag=r4g
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log | cut -c1-220
BPV_OVERLAP=7 python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest_fused.log 2>&1; echo "pytest fused rc=$?" >> gpurun_out/${tag}_pytest_fused.log
tail -6 gpurun_out/${tag}_pytest_fused.log | cut -c1-220
BPV_OVERLAP=7 python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2_fused.json 2> gpurun_out/${tag}_bench.err; echo "bench fused rc=$?"
python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2.json 2>> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_c2_fused.json','gpurun_out/${tag}_bench_c2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
        print({k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
        print('roofline',d['roofline']['frac'],'by_time',d['roofline_by_time'].get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
