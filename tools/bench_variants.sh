# usage: bash tools/bench_variants.sh [variant ...]   (run on the GPU box; variants are .so files under lib/variants)
for v in default "$@"; do
  if [ "$v" = default ]; then unset UKFB_LIB; else export UKFB_LIB=$PWD/slam_pose_estimation_b200/lib/variants/$v.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],4), 'ms')"
done
