cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu22.log 2>&1; tail -5 gpurun_out/pytest_gpu22.log
for lib in "" _prev; do
  CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$lib.so timeout 300 python tests/profile_step.py --pop 64 --evals 5 2>&1 | grep "step ms"
  CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$lib.so timeout 300 python tests/profile_step.py --pop 64 --evals 3 --timing 2>&1 | grep -E "total conv|G1[3-6]|D[01]:c|step ms"
done
