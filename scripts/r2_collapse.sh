#!/bin/bash
# vectorised collapse kernels: parity subset + e2e timing
timeout 400 python -m pytest tests/test_golden.py tests/test_render_gpu.py tests/test_parity_gpu.py -x -q -m gpu -k "golden or opencv or render or small_sequence or band_number or save_png or tile_shards or spread_map or ragged or display or idempot" > gpurun_out/r2_v17_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_v17_tests.log
timeout 200 python bench.py --only --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_v17_e2e.json 2> gpurun_out/r2_v17_e2e.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_v17_e2e.json'))
    e=d['e2e']; print('value %.0f (%.3f ms)  e2e %.0f Mpix/s  %.2f ms  feed %.2f  save %.2f  sha %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['breakdown_ms']['feed_batch_from_host'], e['breakdown_ms']['collapse_and_d2h'], e['mosaic_sha256']))
except Exception as ex: print('parse failed', ex)
PY
