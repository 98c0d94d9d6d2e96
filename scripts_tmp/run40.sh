cd $GRAFT_REPO_ROOT
GLASS_DEBUG_C1_MODE4=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_against or tiny_layerwise_tcgen05" > gpurun_out/pytest_gpu40.log 2>&1; tail -3 gpurun_out/pytest_gpu40.log
GLASS_DEBUG_C1_MODE4=1 timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown40_m4.log 2>&1; grep -E "step ms|total conv|^D0" gpurun_out/breakdown40_m4.log
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown40.log 2>&1; grep -E "step ms|total conv|^D0" gpurun_out/breakdown40.log
