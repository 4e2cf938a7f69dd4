#!/bin/bash
# first GPU pass of the Map2DRender (type 4) path + A/B of the constant-denominator coordinate fast path
TAG=${1:-v15}
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_adapter.py -x -q -m gpu -k "render" > gpurun_out/r2_${TAG}_render_tests.log 2>&1
echo "render pytest rc=$?"; tail -15 gpurun_out/r2_${TAG}_render_tests.log
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "band" > gpurun_out/r2_${TAG}_band_tests.log 2>&1
echo "band pytest rc=$?"; tail -3 gpurun_out/r2_${TAG}_band_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_${TAG}_smoke.log 2>&1
echo "smoke rc=$?"; tail -8 gpurun_out/r2_${TAG}_smoke.log
timeout 300 python bench.py --only --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_${TAG}_mb_only.json 2> gpurun_out/r2_${TAG}_mb_only.err
echo "bench mb rc=$?"
timeout 300 python bench.py --mode weighted --only --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_${TAG}_w_only.json 2> gpurun_out/r2_${TAG}_w_only.err
echo "bench w rc=$?"
python - <<PY
import json
for f in ('mb','w'):
    try:
        d=json.load(open('gpurun_out/r2_${TAG}_%s_only.json' % f))
        print(f, 'value %.0f ms %.3f frac %.3f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['path']['frac'], d['parity']['identical'] if d.get('parity') else None))
        print({k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
    except Exception as e: print(f, 'parse failed', e)
PY
