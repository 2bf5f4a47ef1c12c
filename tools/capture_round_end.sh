# evidence of a build in one GPU call: GPU tests (parity record), smoke, bench lines, ncu launch list, full captures of the
# step path and the rollout kernel (each only after the same command has exited 0 without ncu)
set -x
python -m pytest tests -m gpu -q > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 --burn-in 100 > gpurun_out/bench20.log 2> gpurun_out/bench20.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --burn-in 100 > gpurun_out/ncu_launch.log 2>&1
python tools/steady_steps.py 262144 400 4 > gpurun_out/steady.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"step_kernel_v2|close_kernel" --launch-skip 802 --launch-count 2 -f -o gpurun_out/r02_step_final python tools/steady_steps.py 262144 400 4 > gpurun_out/ncu_step.log 2>&1
python tools/ab_rollout.py cap 65536 > gpurun_out/rollout.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/r02_rollout python tools/ab_rollout.py cap 65536 > gpurun_out/ncu_rollout.log 2>&1
grep -E "passed|failed|rc=" gpurun_out/gputest.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/steady.log gpurun_out/rollout.log; head -c 600 gpurun_out/bench_final.json
