# usage (on the GPU box, from the repo root): bash tools/gpu_profile.sh TAG
# plain runs first (each must exit 0), then the same commands under ncu; reports land in gpurun_out/
TAG=${1:-r01c}
set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-literal --no-orientation > gpurun_out/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ukf_pose_fast -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_pose_fast \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-literal --no-orientation > gpurun_out/ncu_pose_${TAG}.log 2>&1
python tools/bench_c2.py --cases c2 --batches 1048576 --reps 2 > gpurun_out/plain3_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ukf_ori_fast -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_ori_fast \
    python tools/bench_c2.py --cases c2 --batches 1048576 --reps 2 > gpurun_out/ncu_ori_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_ori_${TAG}.log
