#!/bin/bash
# group-size sweep with the TMA pyrDown + launch list of one step
run() { name=$1; shift
  env $ENVV timeout 200 python bench.py --only --steps 10 --warmup 3 --no-cpu --no-e2e "$@" 2>gpurun_out/r2_sweep_$name.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$name value %.0f Mpix/s ms/step %.3f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']), {k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()}, d['roofline']['need_px_fraction'])
" | tee -a gpurun_out/r2_sweep_summary.txt
}
ENVV="M2D_NONE=0" run b500 --batch 500
ENVV="M2D_NONE=0" run b250 --batch 250
ENVV="M2D_NONE=0" run b167 --batch 167
ENVV="M2D_NONE=0" run b125 --batch 125
ENVV="M2D_NONE=0" run b100 --batch 100
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_v11_launches.csv python scripts/prof_run.py --mode multiband --frames 500 --reps 2 > gpurun_out/r2_v11_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 40 -c 34 -o gpurun_out/r2_v11_mb_full python scripts/prof_run.py --mode multiband --frames 500 --reps 2 > gpurun_out/r2_v11_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -o gpurun_out/r2_v11_w_full python scripts/prof_run.py --mode weighted --frames 100 --reps 3 > gpurun_out/r2_v11_ncu3.log 2>&1
echo done
