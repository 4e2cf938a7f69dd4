#!/bin/bash
TAG=${1:-v12}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_${TAG}_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_${TAG}_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_${TAG}_bench.json 2> gpurun_out/r2_${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_${TAG}_bench.json'))
    print('MB value %.0f ms %.3f e2e %.0f (%s) frac %.3f ub %.3f launches %d parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['breakdown_ms'], d['roofline']['path']['frac'], d['roofline']['path']['frac_upper_bound'], d['gpu_launches'], d['parity']['identical']))
    print({k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
    w=d['weighted']; print('W value %.0f ms %.3f e2e %.0f (%s) frac %.3f ub %.3f parity %s' % (w['value'], w['ms_per_step'], w['e2e']['value'], w['e2e']['breakdown_ms'], w['roofline']['path']['frac'], w['roofline']['path']['frac_upper_bound'], w['parity']['identical']))
    print('stream', d.get('stream_latency')); print('cfg3', d.get('cfg3'))
except Exception as e: print('bench parse failed', e)
PY
