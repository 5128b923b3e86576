#!/bin/bash
# A/B helper for kernel work (run under gpurun): the GPU parity tests of the conformer path, then the bench step twice
# with the per-kernel share of the step.  Environment switches of experimental kernels are passed through, e.g.
#     gpurun --timeout 300 -- 'MAFE_HALFWARP_SWEEP=1 bash tools/ab_bench.sh'
timeout 300 python -m pytest tests/test_gpu_features.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step %.3f' % d['ms_per_step'], {k: (round(v,3) if isinstance(v,float) else v) for k,v in d['roofline']['step_share'].items() if k!='note'}, d['clocks'])
"
done
