cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; tail -5 gpurun_out/pytest_gpu23.log
for rep in 1 2; do
for v in new gen prev; do
  export CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200.so; unset GLASS_DEBUG_GENERIC_EPI
  [ $v = gen ] && export GLASS_DEBUG_GENERIC_EPI=1
  [ $v = prev ] && export CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200_prev.so
  echo "== $v"
  timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing 2>&1 | grep -E "total conv|G1[3-6]|D[01]:c|step ms"
done; done
