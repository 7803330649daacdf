#!/bin/bash
# On the GPU box (one gpurun call): parity suite + headline grid for the product build and every experiment build made
# by tools/build_experiments.sh.
for lib in "" gypsum_b200/exp_layout_b.so gypsum_b200/exp_spec_alias.so gypsum_b200/exp_both.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== ${lib:-product build}"
  GB200_LIB=$lib timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
  for i in 1 2; do GB200_LIB=$lib python tools/quick_grid.py 2>&1 | grep -o '"device_ms[^,]*, [^,]*, [^,]*, [^,]*, [^,]*, "correlate_cells_ms": [0-9.]*'; done
done
# tracking variant: config 4 timing with the product build and the fast-angle build
for lib in "" gypsum_b200/exp_fast_angle.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== tracking, ${lib:-product build}"
  [ -n "$lib" ] && GB200_LIB=$lib timeout 600 python -m pytest tests/test_gpu_tracker.py tests/test_gpu_bits.py tests/test_gpu_receiver_flow.py -q -x 2>&1 | tail -2
  GB200_LIB=$lib python - <<'PY' 2>&1 | grep -o '"seconds": [0-9.]*, "channel_ms_per_s": [0-9.]*'
import sys; sys.argv=["x"]; sys.path.insert(0,"tools"); import bench_configs as b; b.tracker_case(32, 60000)
PY
done
