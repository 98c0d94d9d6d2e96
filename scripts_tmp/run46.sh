cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_against or streamed_tap or properties" > gpurun_out/pytest_gpu46.log 2>&1; tail -3 gpurun_out/pytest_gpu46.log
timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing > gpurun_out/breakdown46.log 2>&1; grep -E "step ms|total conv|^D0:c1" gpurun_out/breakdown46.log
