import sys; sys.argv=["x"]; sys.path.insert(0,"tools"); import bench_configs as b; b.tracker_case(32, 20000)
