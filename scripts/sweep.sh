#!/bin/bash
# quick sweep of the group size on the bench workload (prints value and per-kernel avg us)
for mode in multiband weighted; do
for b in 4 8 16 32; do
python bench.py --mode $mode --steps 3 --warmup 3 --no-cpu --no-e2e --batch $b 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$mode batch $b value %.0f Mpix/s ms/step %.2f'%(d['value'],d['ms_per_step']), {k:(v['launches'],v['avg_us']) for k,v in d['roofline']['kernels'].items()})
"
done; done
