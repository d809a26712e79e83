#!/bin/bash
# the driver's own command lines, for the record
tag=${1:-r2n}
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c2_reference_arm.json 2> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c2_steps20.json 2>> gpurun_out/${tag}_bench.err; echo "bench20 rc=$?"
python bench.py --steps 200 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2>> gpurun_out/${tag}_bench.err; echo "bench200 rc=$?"
tail -c 800 gpurun_out/${tag}_bench.err
python - <<PY
import json
for f in ('${tag}_bench_c2_reference_arm','${tag}_bench_c2_steps20','${tag}_bench_c2'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f,'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'steps',d['steps'],d['warmup'], 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind'))
    if 'kernels' in d:
        print('  ',{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()},'F1 frac',round(d['roofline']['frac'],4),'F2 fp64 frac',round(d['roofline_by_time']['frac'],3),'clocks',d['clocks'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
