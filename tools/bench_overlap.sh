# usage (GPU box): bash tools/pdl_test.sh LIB   -- device-timed bench step at several batch sizes, launches overlapped (default),
# overlapped at any grid size (UKFB_OVERLAP_MAX_WAVES=1000) and not overlapped (UKFB_OVERLAP_LAUNCHES=0)
LIB=${1:-slam_pose_estimation_b200/lib/libukfb.so}
for F in 1048576 524288 262144 131072 65536; do
for mode in "default:1:8" "always:1:1000" "off:0:8"; do
  IFS=: read name ov mw <<< "$mode"
  UKFB_OVERLAP_LAUNCHES=$ov UKFB_OVERLAP_MAX_WAVES=$mw UKFB_LIB=$PWD/$LIB timeout 300 python bench.py --filters $F --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-literal --no-orientation 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$F filters, overlap $name: pose', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],4), 'ms', flush=True)"
done
done
