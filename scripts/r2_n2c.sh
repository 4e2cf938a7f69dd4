#!/bin/bash
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_adapter.py -x -q -m gpu -k "multi_device or two_gpus" > gpurun_out/r2_n2c_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_n2c_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --cfg3-frames 0 --cfg4-frames 400 --no-e2e > gpurun_out/r2_n2c_bench.json 2> gpurun_out/r2_n2c_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_n2c_bench.err | cut -c1-400
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_n2c_bench.json'))
    print('N=2 MB value %.0f ms %.3f' % (d['value'], d['ms_per_step']))
    print('cfg4', d.get('cfg4'))
except Exception as e: print('parse failed', e)
PY
