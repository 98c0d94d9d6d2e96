cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu29.log 2>&1; tail -5 gpurun_out/pytest_gpu29.log
timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing > gpurun_out/breakdown29.log 2>&1; grep -E "step ms|total conv|C0:|C5:" gpurun_out/breakdown29.log
for bn in 128 64; do
GLASS_DEBUG_GEMM_BN=$bn timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing > gpurun_out/breakdown29_bn$bn.log 2>&1; grep -E "step ms|total conv|C0:|C5:" gpurun_out/breakdown29_bn$bn.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"from_rgb_fir|fir_down|blur_s2d|layernorm|vecmat_batched" -c 16 -o gpurun_out/hbm29 python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu29.log 2>&1
