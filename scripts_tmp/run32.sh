cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu32.log 2>&1; tail -5 gpurun_out/pytest_gpu32.log
timeout 300 python tests/profile_step.py --pop 64 --evals 10 2>&1 | grep "step ms"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"from_rgb_fir|fir_down" -c 4 -o gpurun_out/hbm32 python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu32.log 2>&1
