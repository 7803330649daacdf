#!/bin/bash
for a in base s300 s700; do
  if [ $a = base ]; then unset GB200_LIB; else export GB200_LIB=$PWD/abl_$a.so; fi
  python - <<PY
import sys, json; sys.argv=["x"]; sys.path.insert(0,"tools")
import bench_configs as b
print("== $a")
b.grid_case("config 2 x 32 blocks", 2046, 1, 41, 32, 50)
b.grid_case("config 3", 4092, 10, 41, 1, 20)
b.grid_case("config 5 shape", 16368, 1, 81, 4, 5)
PY
done 2>&1 | grep -E "==|workload" | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('=='): print(l); continue
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('   ', d['workload'], 'correlate_ms', round(d['correlate_cells_ms'],4))
"
