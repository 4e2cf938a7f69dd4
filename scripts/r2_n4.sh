#!/bin/bash
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_n4_bench.json 2> gpurun_out/r2_n4_bench.err
echo "bench rc=$?"; grep -v Warning gpurun_out/r2_n4_bench.err | tail -2 | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_n4_bench.json'))
    print('N=4 MB value %.0f ms %.3f save %.2f incl %.0f e2e %s parity %s' % (d['value'], d['ms_per_step'], d['save']['ms'], d['save']['value_incl_save'], d['e2e'] and round(d['e2e']['value']), d['parity']['identical']))
    w=d.get('weighted'); print('  W', w and (round(w['value']), w['ms_per_step'], w['parity']['identical']))
    c=d.get('cfg3'); print('  cfg3', c and (round(c['value']), c['ms_per_step'], c.get('mosaic_sha256')))
except Exception as e: print('parse failed', e)
PY
