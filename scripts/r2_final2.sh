#!/bin/bash
# round-2 closing run: full GPU test suite, the driver-shaped bench line, ncu launch lists + --set full captures
TAG=${1:-v16}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_${TAG}_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_${TAG}_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_${TAG}_bench.json 2> gpurun_out/r2_${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_${TAG}_bench.json'))
    print('MB value %.0f ms %.3f e2e %.0f (%.2f ms) frac %.3f ub %.3f launches %d parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['path']['frac'], d['roofline']['path']['frac_upper_bound'], d['gpu_launches'], d['parity']['identical']))
    print({k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
    w=d['weighted']; print('W value %.0f ms %.3f e2e %.0f (%.2f ms) frac %.3f ub %.3f parity %s' % (w['value'], w['ms_per_step'], w['e2e']['value'], w['e2e']['ms_per_step'], w['roofline']['path']['frac'], w['roofline']['path']['frac_upper_bound'], w['parity']['identical']))
    print('stream', {k:(round(v['p50_ms'],3), round(v['p99_ms'],3)) for k,v in d['stream_latency'].items() if isinstance(v, dict)}); print('cfg3', d['cfg3']['value'], d['cfg3']['ms_per_step'], d['cfg3']['save_ms'])
    print('render', d.get('render'))
except Exception as e: print('bench parse failed', e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_${TAG}_launches.csv python scripts/prof_run.py --mode multiband --frames 500 --reps 2 > gpurun_out/r2_${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 42 -c 36 -o gpurun_out/r2_${TAG}_mb_full python scripts/prof_run.py --mode multiband --frames 500 --reps 2 > gpurun_out/r2_${TAG}_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -c 40 -k regex:'mbs_mark|mbs_pull|mbs_warp' -o gpurun_out/r2_${TAG}_pull_full python scripts/prof_run.py --mode multiband --frames 250 --reps 1 --pinned > gpurun_out/r2_${TAG}_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -c 40 -k regex:'rnd_|mosaic_upadd' -o gpurun_out/r2_${TAG}_render_full python scripts/prof_run.py --mode render --frames 20 --reps 2 > gpurun_out/r2_${TAG}_ncu4.log 2>&1
echo ncu done
