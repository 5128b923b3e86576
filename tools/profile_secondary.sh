#!/bin/bash
# One `ncu --set full` launch per secondary kernel (run under gpurun, one GPU); the raw pages are summarised into
# profiles/<round>_secondary_kernels_ncu.txt by tools/summarise_secondary.py.
R=${1:-r02}
cd "$(dirname "$0")/.."
run() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none -k regex:"$rx" -s $skip -c 1 -f -o gpurun_out/${R}_sec_$name "$@" > /dev/null 2>&1
  ncu -i gpurun_out/${R}_sec_$name.ncu-rep --page raw --csv > gpurun_out/${R}_sec_$name.csv 2>/dev/null
  rm -f gpurun_out/${R}_sec_$name.ncu-rep
}
run stft512 "stft512_kernel" 3 python tools/bench_stft.py
run stftn16_20 "stftn16_kernel" 3 python tools/bench_ds2.py
run stftn16_25 "stftn16_kernel" 3 python tools/bench_stft.py --n-fft 400 --hop 160
run scalar_norm "scalar_norm_apply" 2 python tools/bench_ds2.py
run fbank400 "fbank400_kernel" 3 python tools/bench_features.py
run db_clamp "db_clamp_kernel" 3 python tools/bench_features.py
run dct "dct_" 3 python tools/bench_features.py --mfcc
run front2048 "front2048_kernel" 3 python tools/bench_features.py --fastspeech2
run cmvn_stats "cmvn_stats" 2 python tools/bench_cfg4.py --steps 3
run cmvn_apply "cmvn_apply" 2 python tools/bench_cfg4.py --steps 3
ls gpurun_out/${R}_sec_*.csv | wc -l
