#!/bin/bash
# round 2, 8 GPUs: the N=8 bench line (weak cfg2 both modes + sharded save + in-run parity, cfg3 strong, cfg4)
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_n8_bench.json 2> gpurun_out/r2_n8_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_n8_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_n8_bench.json'))
    print('N=8 MB value %.0f ms %.3f save %.2f ms incl %.0f e2e %s parity %s' % (d['value'], d['ms_per_step'], d['save']['ms'], d['save']['value_incl_save'], d['e2e'] and round(d['e2e']['value']), d['parity']['identical']))
    print(d['run']['parallelism'])
    w=d.get('weighted'); print('W', w and (round(w['value']), w['ms_per_step'], w['save']['ms'], w['parity']['identical'], w.get('e2e')))
    c=d.get('cfg3'); print('cfg3', c and (round(c['value']), c['ms_per_step'], c['save']['ms'], c.get('mosaic_sha256'), c['parallelism']))
    c=d.get('cfg4'); print('cfg4', c)
except Exception as e: print('parse failed', e)
PY
