#!/bin/bash
# round 2, 2 GPUs: sharded parity tests over NCCL, then the N=2 bench line (weak cfg2 both modes + cfg3 strong + sharded save + in-run parity)
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > gpurun_out/r2_n2_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_n2_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_n2_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_n2_bench.json'))
    print('N=2 MB value %.0f ms %.3f save %.2f ms incl %.0f e2e %s parity %s' % (d['value'], d['ms_per_step'], d['save']['ms'], d['save']['value_incl_save'], d['e2e'] and round(d['e2e']['value']), d['parity']))
    w=d.get('weighted'); print('W', w and (round(w['value']), w['ms_per_step'], w['save']['ms'], w['parity']))
    c=d.get('cfg3'); print('cfg3', c and (round(c['value']), c['ms_per_step'], c['save'], c.get('mosaic_sha256'), c['parallelism']))
except Exception as e: print('parse failed', e)
PY
