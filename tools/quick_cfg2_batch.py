"""Config 2 with 32 / 128 / 512 blocks per call (needs GB200_SPEC_BUDGET_MB large enough to keep each call one launch pair)."""
import sys

sys.argv = ["x"]
sys.path.insert(0, "tools")
import bench_configs as b  # noqa: E402

for nb, reps in ((32, 100), (64, 60), (96, 40), (128, 30), (256, 15)):
    b.grid_case(f"config 2 x {nb} blocks", 2046, 1, 41, nb, reps)
