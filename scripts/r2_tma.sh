#!/bin/bash
# round 2: TMA-staged pyrDown — parity first (tight timeout: a wrong expect_tx would hang on the mbarrier), then A/B timing
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "weights_first or small_sequence or benchmarked" > gpurun_out/r2_tma_tests.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -6 gpurun_out/r2_tma_tests.log
if [ $rc -ne 0 ]; then exit 1; fi
run() { name=$1; shift
  env $ENVV timeout 200 python bench.py --only --steps 10 --warmup 3 --no-cpu --no-e2e "$@" 2>gpurun_out/r2_tma_$name.err | tee gpurun_out/r2_tma_$name.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$name value %.0f Mpix/s ms/step %.3f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']), {k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
" | tee -a gpurun_out/r2_tma_summary.txt
}
ENVV="M2D_TMA=1" run tma1
ENVV="M2D_TMA=0" run tma0
ENVV="M2D_TMA=1" run tma1_w --mode weighted
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tma_alltests.log 2>&1
echo "all tests rc=$?"; tail -4 gpurun_out/r2_tma_alltests.log
timeout 300 python - <<PY
import sys; sys.path.insert(0, '.')
import bench, argparse
import pi_slam_fusion_b200.map2d as m2d
a = argparse.Namespace()
for name, typ in (("multiband", 3), ("weighted", 1)):
    print(name, bench.stream_latency(a, m2d, typ, 0, 2000))
PY
