cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_against or tiny_layerwise_tcgen05 or properties" > gpurun_out/pytest_gpu45.log 2>&1; tail -3 gpurun_out/pytest_gpu45.log
for v in "" _parked ""; do
  export CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$v.so
  timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing > gpurun_out/breakdown45$v.log 2>&1; grep -E "step ms|total conv|^G16|^D0:c1|^D1:c1|^G8|^G13" gpurun_out/breakdown45$v.log
done
