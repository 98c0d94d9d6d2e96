cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu24.log 2>&1; tail -5 gpurun_out/pytest_gpu24.log
GLASS_DEBUG_SPEC_LOG=1 timeout 300 python tests/profile_step.py --pop 64 --evals 1 2>&1 | grep "conv_tc:" | grep -v "H=1 " > gpurun_out/spec_log.txt
for rep in 1 2; do
for v in new gen; do
  unset GLASS_DEBUG_GENERIC_EPI
  [ $v = gen ] && export GLASS_DEBUG_GENERIC_EPI=1
  echo "== $v"
  timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing 2>&1 | grep -E "total conv|G1[0-6]|D[0-3]:|step ms"
done; done
