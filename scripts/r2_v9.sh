#!/bin/bash
# round 2: parity of the culled weights-first pipeline, then a first timing
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/r2_v9_tests.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_v9_tests.log
run() { # name, bench args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 240 python bench.py --mode multiband --steps 5 --warmup 3 --no-cpu --no-e2e "$@" 2>gpurun_out/r2_v9_$name.err | tee gpurun_out/r2_v9_$name.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$name value %.0f Mpix/s ms/step %.3f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']), {k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
" | tee -a gpurun_out/r2_v9_summary.txt
}
ENVV="M2D_NONE=0" run default
ENVV="M2D_NONE=0" run b125 --batch 125
ENVV="M2D_NONE=0" run b250 --batch 250
ENVV="M2D_WCULL=0" run nocull_b125 --batch 125
ENVV="M2D_SPARSE=0" run dense
