#!/bin/bash
# pull mode (pinned host frames): parity tests + e2e A/B
TAG=${1:-v16}
timeout 600 python -m pytest tests/test_pull_gpu.py -x -q -m gpu > gpurun_out/r2_${TAG}_pull_tests.log 2>&1
echo "pull pytest rc=$?"; tail -15 gpurun_out/r2_${TAG}_pull_tests.log
for mode in multiband weighted; do
  for zc in 1 0; do
    M2D_ZEROCOPY=$zc timeout 300 python bench.py --mode $mode --only --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_${TAG}_e2e_${mode}_zc${zc}.json 2> gpurun_out/r2_${TAG}_e2e_${mode}_zc${zc}.err
    echo "$mode zc=$zc rc=$?"
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_${TAG}_e2e_${mode}_zc${zc}.json'))
    e=d['e2e']; print('  value %.0f (%.3f ms)  e2e %.0f Mpix/s  %.2f ms  feed %.2f  save %.2f  sha %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['breakdown_ms']['feed_batch_from_host'], e['breakdown_ms']['collapse_and_d2h'], e['mosaic_sha256']))
except Exception as ex: print('  parse failed', ex)
PY
  done
done
