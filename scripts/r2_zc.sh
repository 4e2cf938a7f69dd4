#!/bin/bash
# A/B: e2e with H2D staging copies vs kernels sampling the pinned host frames in place over PCIe
for mode in multiband weighted; do
  for zc in "" "--zero-copy"; do
    timeout 300 python bench.py --mode $mode --only --steps 5 --warmup 3 --no-cpu $zc > gpurun_out/r2_zc_${mode}_${zc:2}.json 2> gpurun_out/r2_zc_${mode}_${zc:2}.err
    echo "$mode '$zc' rc=$?"
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_zc_${mode}_${zc:2}.json'))
    e=d['e2e']; print('  e2e %.0f Mpix/s  %.2f ms  feed %.2f  save %.2f  sha %s' % (e['value'], e['ms_per_step'], e['breakdown_ms']['feed_batch_from_host'], e['breakdown_ms']['collapse_and_d2h'], e['mosaic_sha256']))
except Exception as ex: print('  parse failed', ex)
PY
  done
done
