#!/bin/bash
# round 2, 2 GPUs: multi-device handle + adapter on two GPUs + sharded tests, smoke, then the N=2 bench line again
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_n2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_n2b_smoke.log
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_adapter.py tests/test_parity_gpu.py -x -q -m gpu -k "sharded or multi_device or adapter or weights_first or small_sequence or benchmarked" > gpurun_out/r2_n2b_tests.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_n2b_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_n2b_bench.json 2> gpurun_out/r2_n2b_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_n2b_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_n2b_bench.json'))
    print('N=2 MB value %.0f ms %.3f save %.2f ms incl %.0f e2e %s parity %s' % (d['value'], d['ms_per_step'], d['save']['ms'], d['save']['value_incl_save'], d['e2e'] and round(d['e2e']['value']), d['parity']['identical']))
    w=d.get('weighted'); print('W', w and (round(w['value']), w['ms_per_step'], w['save']['ms'], w['parity']['identical']))
    c=d.get('cfg3'); print('cfg3', c and (round(c['value']), c['ms_per_step'], c['save']['ms'], c.get('mosaic_sha256')))
except Exception as e: print('parse failed', e)
PY
