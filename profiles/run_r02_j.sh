cd $GRAFT_REPO_ROOT
TAG=${1:-r02j}
# A/B of the fused-FIR down-convs inside one run: 3 interleaved rounds x 20 evaluations per variant
for round in 1 2 3; do
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/both fused    /"
  CLIPGLASS_LIB=$GRAFT_REPO_ROOT/clip_glass_b200/libclipglass_b200_dbg.so GLASS_DEBUG_NOFUSED64=1 timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/D0 fused only /"
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 --flags 512 2>&1 | grep "step ms" | sed "s/^/none fused    /"
done > gpurun_out/ab_$TAG.log
cat gpurun_out/ab_$TAG.log | cut -c1-60
