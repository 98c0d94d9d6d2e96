cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu42.log 2>&1; tail -15 gpurun_out/pytest_gpu42.log
timeout 300 python tests/profile_step.py --pop 64 --evals 8 2>&1 | grep "step ms"
timeout 300 python tests/profile_step.py --pop 64 --evals 8 --flags 64 2>&1 | grep "step ms"
timeout 300 python tests/profile_step.py --pop 64 --evals 8 2>&1 | grep "step ms"
timeout 300 python tests/profile_step.py --pop 64 --evals 8 --flags 64 2>&1 | grep "step ms"
