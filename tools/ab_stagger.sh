for ab in "600 150" "0 0" "2200 550" "1100 275" "3300 800" "2200 0" "0 550"; do
  set -- $ab
  echo "stagger $1 $2: $(GB200_STAGGER_A=$1 GB200_STAGGER_B=$2 python tools/quick_grid.py 2>&1 | grep -o '"device_ms[^,]*, [^,]*, [^,]*, [^,]*, [^,]*, "correlate_cells_ms": [0-9.]*')"
done
