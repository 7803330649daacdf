#!/bin/bash
# One gpurun call: GPU tests + secondary configs (+ optional profiles).  usage: gpu_round.sh TAG [profile]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python tools/bench_configs.py 2>&1 | tee gpurun_out/configs_${TAG}.jsonl | cut -c1-600
GB200_DETECT_FUSED=0 timeout 300 python - <<'PY' 2>&1 | tee -a gpurun_out/configs_${TAG}.jsonl | cut -c1-400
import sys; sys.argv=["x"]; sys.path.insert(0,"tools"); import bench_configs as b; b.detector_case()
PY
if [ "${2:-}" = "profile" ]; then
  bash tools/gpu_profile.sh ${TAG}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_track_channels -s 1 -c 1 \
      -f -o gpurun_out/prof_track_${TAG} python tools/profile_tracker.py > gpurun_out/ncu_track_${TAG}.log 2>&1
else
  python bench.py | tee gpurun_out/bench_${TAG}.json | cut -c1-1500
fi
