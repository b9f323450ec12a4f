# usage (GPU box): bash tools/gpu_variants.sh TAG lib1.so lib2.so ...   -- pose C4 + orientation C2 for each build (two
# passes), then the GPU parity tests of the LAST build listed; output in gpurun_out/variants_TAG.txt
TAG=$1; shift
OUT=gpurun_out/variants_${TAG}.txt
: > $OUT
bash tools/bench_libs.sh "$@" >> $OUT 2>&1
LAST=${@: -1}
UKFB_LIB=$PWD/$LAST python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_ref_pin.py tests/test_events.py tests/test_gate.py -m gpu -x -q 2>&1 | tail -3 >> $OUT
cat $OUT
