#!/bin/bash
# tuning sweep over lanes-per-problem / in-slab buffering for the (8,4) kernels (run on the GPU box)
for dt in f64 f32; do for L in 4 8; do for NB in 1 2; do
  echo "== dtype=$dt L=$L NBUF=$NB"
  IDOC_L=$L IDOC_NBUF=$NB python bench.py --dtype $dt --steps 5 --warmup 2 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); k=d['roofline']['kernels']
        print('  ms/step %.3f  solves/s %.3g  frac %.3f | '%(d['ms_per_step'],d['value'],d['roofline']['frac'])+' '.join('%s %.3f'%(n[4:],v['mean_ms']) for n,v in k.items()))
    else: print(line.rstrip())
"
done; done; done
