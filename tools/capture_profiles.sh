set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench20.log 2> gpurun_out/bench20.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --burn-in 100 > gpurun_out/ncu_launch.log 2>&1
python tools/steady_steps.py 262144 400 4 > gpurun_out/steady.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"step_kernel_v2|close_kernel" --launch-skip 802 --launch-count 2 -o gpurun_out/r02_step_final python tools/steady_steps.py 262144 400 4 > gpurun_out/ncu_step.log 2>&1
python tools/ab_rollout.py cap 65536 > gpurun_out/rollout.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/r02_rollout python tools/ab_rollout.py cap 65536 > gpurun_out/ncu_rollout.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:curiosity_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r02_curiosity python tools/ab_curiosity.py 262144 4 > gpurun_out/ncu_cur.log 2>&1
tail -2 gpurun_out/steady.log gpurun_out/rollout.log gpurun_out/ncu_step.log gpurun_out/ncu_rollout.log gpurun_out/ncu_cur.log
