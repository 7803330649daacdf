#!/bin/bash
# On the GPU box (one gpurun call): parity suite + headline grid for the product build and every experiment build made
# by tools/build_experiments.sh.
for lib in "" gypsum_b200/exp_layout_b.so gypsum_b200/exp_spec_alias.so gypsum_b200/exp_both.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "== ${lib:-product build}"
  GB200_LIB=$lib timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
  for i in 1 2; do GB200_LIB=$lib python tools/quick_grid.py 2>&1 | grep -o '"device_ms[^,]*, [^,]*, [^,]*, [^,]*, [^,]*, "correlate_cells_ms": [0-9.]*'; done
done
