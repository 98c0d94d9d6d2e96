set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; tail -3 gpurun_out/pytest_gpu21.log
for lib in "" _cta3 _cta2; do
  CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$lib.so timeout 300 python tests/profile_step.py --pop 64 --evals 5 2>&1 | grep "step ms"
  CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$lib.so timeout 300 python tests/profile_step.py --pop 64 --evals 3 --timing 2>&1 | grep -E "total conv|G16|D0:c0|step ms"
done
GLASS_DEBUG_SPLIT_FRGB=1 timeout 300 python tests/profile_step.py --pop 64 --evals 5 2>&1 | grep "step ms"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir|from_rgb" -c 24 --csv --log-file gpurun_out/fir21.csv python tests/profile_step.py --pop 64 --evals 2 > /dev/null 2>&1
grep -E "fir|from_rgb" gpurun_out/fir21.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | tail -12
