cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu36.log 2>&1; tail -15 gpurun_out/pytest_gpu36.log
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown36.log 2>&1; grep -E "step ms|total conv|^D[0-7]:c1" gpurun_out/breakdown36.log
timeout 300 python tests/profile_step.py --pop 64 --evals 8 2>&1 | grep "step ms"
