#!/bin/bash
# multi-GPU session (short): NCCL parity tests on real kernels, device-placement test, the N-GPU bench line
tag=${1:-r2z}; n=${2:-2}
mkdir -p gpurun_out
[ -z "$3" ] && python -m pytest tests/test_dist_gpu.py tests/test_round2_gpu.py -m gpu -q --maxfail=10 -k "sharded or second_gpu" > gpurun_out/${tag}_pytest_multi.log 2>&1; tail -4 gpurun_out/${tag}_pytest_multi.log | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 60 --warmup 5 \
   > gpurun_out/${tag}_bench_c2_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${tag}_bench_${n}gpu.err
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_c2_${n}gpu.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'F1 frac',d['roofline']['frac'],'launches',d['gpu_launches'])
        for k,v in d.get('other_shapes',{}).items(): print(k,{a:b for a,b in v.items() if a not in ('desc','cpu')})
    except Exception as e: print(f, 'ERR', e)
PY
