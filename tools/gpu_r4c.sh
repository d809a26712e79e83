#!/bin/bash
# r4c: xcorr with register-resident coarse values (no cv[] array, 5 CTAs per SM) + cheaper operand build, Welch constants: parity, bench, A/B, FFMA2 probe
tag=r4c
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=25 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log | cut -c1-220
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
BPV_XC_ONE=0 python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2_xc_cv.json 2>> gpurun_out/${tag}_bench.err; echo "bench2 rc=$?"
BPV_OVERLAP=1 python bench.py --steps 100 --warmup 5 --no-cpu --no-other > gpurun_out/${tag}_bench_c2_serial_xc.json 2>> gpurun_out/${tag}_bench.err; echo "bench3 rc=$?"
tools/bin/ffma2_probe > gpurun_out/${tag}_ffma2_probe.txt 2>&1; cat gpurun_out/${tag}_ffma2_probe.txt
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_c2.json','gpurun_out/${tag}_bench_c2_xc_cv.json','gpurun_out/${tag}_bench_c2_serial_xc.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
    print({k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
    print('roofline',d['roofline']['frac'],'by_time',d['roofline_by_time'].get('frac'))
    for k,v in d.get('other_shapes',{}).items(): print(k,v['ms_per_step'])
PY
