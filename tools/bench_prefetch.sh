# usage (GPU box): bash tools/bench_prefetch.sh LIB "TILES..." "BYTES..."  -- next-wave prefetch distance / granularity sweep (pose C4)
LIB=${1:-slam_pose_estimation_b200/lib/libukfb.so}
TILES=${2:-0 auto 592}
BYTES=${3:-128}
for b in $BYTES; do for t in $TILES; do
  export UKFB_PREFETCH_BYTES=$b
  if [ $t = auto ]; then unset UKFB_PREFETCH_TILES; else export UKFB_PREFETCH_TILES=$t; fi
  UKFB_LIB=$PWD/$LIB python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-literal --no-orientation 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiles=$t bytes=$b pose', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],4), 'ms', flush=True)"
done; done
