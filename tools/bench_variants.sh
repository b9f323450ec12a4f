# usage (GPU box): bash tools/bench_variants.sh   -- device-timed pose C4 step and orientation C2 tick for UKFB_FAST_WPB = 1, 2, 4
for w in 1 2 4; do
  export UKFB_FAST_WPB=$w
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-literal 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('WPB=$w pose', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],4), 'ms; ori C2', round(d['orientation_c2']['value']/1e6,1), 'M ticks/s')"
done
