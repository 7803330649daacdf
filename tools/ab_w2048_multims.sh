#!/bin/bash
# A/B: multi-millisecond non-coherent grids on the one-warp-per-transform kernel (default: 8 warps / CTA) vs the warp-pair
# kernel (GB200_W2048=0)
for w in 12 0; do
  echo "GB200_W2048=$w: $(GB200_W2048=$w python - <<'PY' 2>&1 | grep -o '"workload[^}]*' | cut -c1-330
import sys; sys.argv=["x"]; sys.path.insert(0,"tools"); import bench_configs as b
b.grid_case("config 3: 32x41x10 ms @ 4.092 Msps", 4092, 10, 41, 1, 30)
b.grid_case("32x41x10 ms @ 2.046 Msps", 2046, 10, 41, 1, 30)
PY
)"
done
