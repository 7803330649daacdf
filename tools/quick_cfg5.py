"""Config-5 shape (32x81 @ 16.368 Msps), 48 blocks per call: device ms and per-kernel ms.  A/B aid: GB200_SPEC_BUDGET_MB=..."""
import sys

sys.argv = ["x"]
sys.path.insert(0, "tools")
import bench_configs as b  # noqa: E402

b.grid_case("config 5 shape: 32x81 @ 16.368 Msps, 48 blocks", 16368, 1, 81, 48, 5)
