set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_b.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ukf_pose_fast -s 3 -c 1 -f -o gpurun_out/prof_r01_fast_v2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fast2.log 2>&1
tail -2 gpurun_out/ncu_fast2.log
