# refresh the step path's evidence after a change to its sources: GPU tests, bench line, full ncu capture of one steady-state step
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python tools/steady_steps.py 262144 400 4 > gpurun_out/steady.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"step_kernel_v2|close_kernel" --launch-skip 802 --launch-count 2 -f -o gpurun_out/r02_step_final python tools/steady_steps.py 262144 400 4 > gpurun_out/ncu_step.log 2>&1
grep -E "passed|failed|rc=" gpurun_out/gputest.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/steady.log
