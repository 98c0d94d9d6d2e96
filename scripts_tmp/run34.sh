cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu34.log 2>&1; tail -15 gpurun_out/pytest_gpu34.log
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown34.log 2>&1; grep -E "step ms|total conv|^D0|^D1:c0" gpurun_out/breakdown34.log
GLASS_DEBUG_C1_I8=0 timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown34_off.log 2>&1; grep -E "step ms|total conv|^D0|^D1:c0" gpurun_out/breakdown34_off.log
