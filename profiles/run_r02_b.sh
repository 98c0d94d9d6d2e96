# round 2, call B: zero-block skipping of the exact polyphase forms: parity, then A/B of the exact/folded thresholds
cd $GRAFT_REPO_ROOT
TAG=${1:-r02b}
timeout 900 python -m pytest tests -m gpu -x -q -k "exact or resampling or multi_seed or tiny_layerwise_tcgen05 or benchmarked" > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tests/profile_step.py --pop 64 --evals 8 > gpurun_out/step_$TAG.log 2>&1; grep "step ms" gpurun_out/step_$TAG.log
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "total conv" gpurun_out/breakdown_$TAG.log
export CLIPGLASS_LIB=$GRAFT_REPO_ROOT/clip_glass_b200/libclipglass_b200_dbg.so
for g in "512,64" "512,32" "512,16" "256,64" "128,64" "64,64"; do
  echo "GEXACT=$g"; GLASS_DEBUG_GEXACT=$g timeout 300 python tests/profile_step.py --pop 64 --evals 8 2>&1 | grep "step ms"
done
for d in 64 32; do
  echo "DEXACT=$d"; GLASS_DEBUG_DEXACT=$d timeout 300 python tests/profile_step.py --pop 64 --evals 8 2>&1 | grep "step ms"
done
GLASS_DEBUG_GEXACT="256,32" timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_${TAG}_g256.log 2>&1
