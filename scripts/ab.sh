#!/bin/bash
for f in 0 1; do
M2D_FUSED=$f python bench.py --mode multiband --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('fused $f value %.0f Mpix/s ms/step %.3f'%(d['value'],d['ms_per_step']), {k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
"
done
