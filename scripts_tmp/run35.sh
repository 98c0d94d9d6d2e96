cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu35.log 2>&1; tail -15 gpurun_out/pytest_gpu35.log
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown35.log 2>&1; grep -E "step ms|total conv|^D0|^D1:c0" gpurun_out/breakdown35.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 68 -c 1 -o gpurun_out/conv35_c1 python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu35.log 2>&1
