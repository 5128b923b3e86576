#!/bin/bash
# Evidence pass of a round (run under gpurun, one GPU): ncu launch list of the bench step, `--set full` captures of every
# kernel of the step (read them here with `ncu -i ... --page details|raw|source`, summarise into profiles/), sanitizers.
R=${1:-r02}
# 1) launch list of the default bench command (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fbank512|frame_sum|cmvn_utt|tile_prepare|fma_peak" -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --sustain-s 0 --oracle-utts 0 > gpurun_out/launch_bench.log 2>&1
# 2) full capture of the step's kernels at the full bench workload (tile records + main kernel = 2 launches per step with the
#    frame-mean sums fused; SKIP=6 COUNT=3 with MAFE_NO_FUSED_FRAMESUM=1: pre-pass, tile records, main kernel); the warm-up's are skipped
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"fbank512_v6|frame_sum|tile_prepare|cmvn_utt_apply" -s ${SKIP:-4} -c ${COUNT:-2} -f -o gpurun_out/${R}_step python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --sustain-s 0 --oracle-utts 0 > gpurun_out/full_bench.log 2>&1
# 3) the FP32 lane rate of scalar vs packed instructions
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fp32x2_peak tools/fp32x2_peak.cu 2>/dev/null && ./gpurun_out/fp32x2_peak > gpurun_out/${R}_fp32x2_peak.txt 2>&1
# 4) sanitizers
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/san_mem.log 2>&1; tail -3 gpurun_out/san_mem.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/san_race.log 2>&1; tail -3 gpurun_out/san_race.log
ls -la gpurun_out/ | tail -8
