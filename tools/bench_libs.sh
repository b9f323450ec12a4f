# usage (GPU box): bash tools/bench_libs.sh [lib ...]   -- device-timed pose C4 step and orientation C2 tick for several
# builds of the library (slam_pose_estimation_b200/lib/variants/*.so by default; UKFB_LIB selects the build, see _build.py)
LIBS=${@:-slam_pose_estimation_b200/lib/variants/*.so}
for rep in 1 2; do
for lib in $LIBS; do
  UKFB_LIB=$PWD/$lib python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-literal 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib pose', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],4), 'ms; ori C2', round(d['orientation_c2']['value']/1e6,1), 'M ticks/s', d['clocks']['sm_mhz'], flush=True)"
done
done
