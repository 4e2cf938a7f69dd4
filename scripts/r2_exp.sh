#!/bin/bash
# round 2, first GPU call: parity + timing of the opt-in weight-stage variants left unmeasured by round 1
set -x
M2D_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "experimental or weights_first" > gpurun_out/r2_exp_tests.log 2>&1
tail -3 gpurun_out/r2_exp_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --mode multiband --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2_exp_$name.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$name value %.0f Mpix/s ms/step %.3f'%(d['value'],d['ms_per_step']), {k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
" | tee -a gpurun_out/r2_exp_summary.txt
}
run base M2D_NONE=0
run wfused M2D_WFUSED=1
run wlean M2D_WLEAN=1
run wfused_wlean M2D_WFUSED=1 M2D_WLEAN=1
run dcull M2D_DCULL=1
run all3 M2D_WFUSED=1 M2D_WLEAN=1 M2D_DCULL=1
