#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same command, one full capture of the
# dominant kernel.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 20 --warmup 3 --cpu-blocks 1 > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_correlate -s 10 -c 2 \
    -f -o gpurun_out/prof_corr_${TAG} python bench.py --steps 6 --warmup 3 --cpu-blocks 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_doppler_spectra -s 10 -c 1 \
    -f -o gpurun_out/prof_spec_${TAG} python bench.py --steps 6 --warmup 3 --cpu-blocks 1 >> gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
