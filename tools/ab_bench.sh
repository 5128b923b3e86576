#!/bin/bash
# A/B helper for kernel work (run under gpurun): the GPU parity tests of the conformer path, then the bench step twice
# with the per-kernel share of the step.  Environment switches of experimental kernels are passed through, e.g.
#     gpurun --timeout 300 -- 'MAFE_FBANK_V3=1 bash tools/ab_bench.sh'
# With an argument: that libmafe.so is measured instead of the in-tree one (tools/ab_bench.sh scratch/libmafe_prev.so).
cd "$(dirname "$0")/.."
if [ -n "$1" ]; then cp mindaudio_b200/libmafe.so /tmp/_ab_keep.so; cp "$1" mindaudio_b200/libmafe.so; fi
timeout 300 python -m pytest tests/test_gpu_features.py -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do
timeout 120 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5 --sustain-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['step']['kernels_ms']
print('ms_per_step %.3f' % d['ms_per_step'], {a: round(b,3) for a,b in k.items()}, 'oracle', d['oracle_check']['ok'], '%.2e' % d['oracle_check']['max_mixed_err_logmel'], d['clocks']['sm_mhz'])
"
done
if [ -n "$1" ]; then cp /tmp/_ab_keep.so mindaudio_b200/libmafe.so; fi
