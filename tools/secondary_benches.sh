#!/bin/bash
# CUDA-event timings of the secondary paths (run under gpurun, one GPU); one JSON line per tool into gpurun_out/<round>_secondary.jsonl
R=${1:-r02}
cd "$(dirname "$0")/.."
out=gpurun_out/${R}_secondary.jsonl
: > $out
for cmd in "tools/bench_stft.py" "tools/bench_stft.py --n-fft 400 --hop 160" "tools/bench_ds2.py" "tools/bench_features.py" "tools/bench_features.py --mfcc" "tools/bench_features.py --fastspeech2" "tools/bench_cfg4.py"; do
  timeout 200 python $cmd 2>/dev/null | tail -1 >> $out
done
wc -l $out
