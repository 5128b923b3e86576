#!/bin/bash
# Evidence pass of a round (run under gpurun, one GPU): ncu launch list of the bench step, one `--set full` capture of the
# dominant kernel (read it here with `ncu -i ... --page details|raw|source`, summarise into profiles/), sanitizer runs.
# Rename the r01_ prefixes per round.
# 1) launch list of the default bench command (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fbank512|frame_sum|cmvn_utt|fma_peak" -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/launch_bench.log 2>&1
# 2) full capture of the dominant kernel at the full bench workload
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fbank512_v3 -s 2 -c 1 -f -o gpurun_out/r01_fbank512_v3 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/full_bench.log 2>&1
# 3) sanitizer
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/san_mem.log 2>&1; tail -3 gpurun_out/san_mem.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/san_race.log 2>&1; tail -3 gpurun_out/san_race.log
ls -la gpurun_out/ | tail -8
