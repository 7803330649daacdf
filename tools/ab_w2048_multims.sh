#!/bin/bash
# A/B: multi-millisecond grids on the warp-pair kernel (default) vs the one-warp-per-transform kernel with 8 warps / CTA
for w in 12 8; do
  echo "GB200_W2048=$w: $(GB200_W2048=$w python - <<'PY' 2>&1 | grep -o '"workload[^}]*' | cut -c1-330
import sys; sys.argv=["x"]; sys.path.insert(0,"tools"); import bench_configs as b
b.grid_case("config 3: 32x41x10 ms @ 4.092 Msps", 4092, 10, 41, 1, 30)
b.grid_case("32x41x10 ms @ 2.046 Msps", 2046, 10, 41, 1, 30)
PY
)"
done
