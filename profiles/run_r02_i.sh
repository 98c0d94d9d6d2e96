cd $GRAFT_REPO_ROOT
TAG=${1:-r02i}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_fir or multi_seed or benchmarked" > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tests/profile_step.py --pop 64 --evals 14 > gpurun_out/step_$TAG.log 2>&1; grep "step ms" gpurun_out/step_$TAG.log
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "total conv|^D[01]:" gpurun_out/breakdown_$TAG.log
exit 0
GLASS_DEBUG_FUSED64=1 timeout 300 python tests/profile_step.py --pop 64 --evals 14 > gpurun_out/step_${TAG}_f64.log 2>&1; grep "step ms" gpurun_out/step_${TAG}_f64.log
GLASS_DEBUG_FUSED64=1 timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_${TAG}_f64.log 2>&1; grep -E "total conv|^D[01]:" gpurun_out/breakdown_${TAG}_f64.log
