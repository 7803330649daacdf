#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same command, one full capture of each hot kernel.
# Outputs land in gpurun_out/.   usage: gpu_profile.sh TAG
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}.json
SMALL="--steps 2 --warmup 1 --calls-per-step 2 --cpu-blocks 1 --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py $SMALL > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_correlate_w2048 -s 10 -c 2 \
    -f -o gpurun_out/prof_corr_${TAG} python bench.py $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_doppler_spectra -s 10 -c 1 \
    -f -o gpurun_out/prof_spec_${TAG} python bench.py $SMALL >> gpurun_out/ncu_full_${TAG}.log 2>&1
# the same correlate launch with the cache state the preceding doppler_spectra launch left (no flush between kernels)
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:k_correlate_w2048 -s 10 -c 2 \
    -f -o gpurun_out/prof_corr_nocc_${TAG} python bench.py $SMALL >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
