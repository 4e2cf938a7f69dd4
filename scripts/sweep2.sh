#!/bin/bash
for ctx in 3 4 6; do for b in 16 32; do
M2D_CTX=$ctx python bench.py --mode multiband --steps 4 --warmup 3 --no-cpu --no-e2e --batch $b 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ctx $ctx batch $b value %.0f Mpix/s ms/step %.2f'%(d['value'],d['ms_per_step']))
"
done; done
M2D_CTX=4 python bench.py --mode weighted --steps 4 --warmup 3 --no-cpu --batch 32 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('weighted ctx 4 value %.0f ms/step %.2f e2e %.0f (%.1f ms)'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
"
