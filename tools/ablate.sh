#!/bin/bash
# Needs the experiment builds abl_a1.so .. abl_a3.so in the repo root: kernels.cu / warp_fft.cuh compiled with
# -DGB_ABLATE=1..3 around the hooks described in profiles/ablation_r1.md (the hooks are not kept in the product source).
# timing-only ablations of correlate_cells (results are wrong by construction): a1 = no pair exchange / barriers /
# recombination twiddles, a2 = a1 + no replica-spectrum product, a3 = a2 + no tw1 table loads
for a in base a1 a2 a3; do
  if [ $a = base ]; then unset GB200_LIB; else export GB200_LIB=$PWD/abl_$a.so; fi
  python - <<PY
import sys, json; sys.argv=["x"]; sys.path.insert(0,"tools")
import bench_configs as b
print("== $a")
b.grid_case("config 2 x 32 blocks", 2046, 1, 41, 32, 50)
b.grid_case("config 3", 4092, 10, 41, 1, 20)
PY
done 2>&1 | grep -E "==|workload" | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('=='): print(l); continue
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('   ', d['workload'], 'correlate_ms', round(d['correlate_cells_ms'],4), 'spectra_ms', round(d['doppler_spectra_ms'],4))
"
