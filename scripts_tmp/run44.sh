cd $GRAFT_REPO_ROOT
for v in "" _c3; do
  export CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$v.so
  timeout 300 python tests/profile_step.py --pop 64 --evals 6 --timing > gpurun_out/breakdown44$v.log 2>&1; grep -E "step ms|total conv|^G16|^D0:c0" gpurun_out/breakdown44$v.log
done
