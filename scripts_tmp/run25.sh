cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu25.log 2>&1; tail -5 gpurun_out/pytest_gpu25.log
for v in "" _mb2 _prev; do
  export CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$v.so
  echo "== $v"
  timeout 300 python tests/profile_step.py --pop 64 --evals 6 2>&1 | grep -E "step ms"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir|from_rgb|final_cos" -s 9 -c 9 --csv --log-file gpurun_out/fir25$v.csv python tests/profile_step.py --pop 64 --evals 2 > /dev/null 2>&1
  grep -E "fir|from_rgb|final_cos" gpurun_out/fir25$v.csv | awk -F'","' '{print substr($5,1,40), $NF}'
done
unset CLIPGLASS_LIB
python bench.py --steps 10 --warmup 3 > gpurun_out/bench25.json 2> gpurun_out/bench25.err; cat gpurun_out/bench25.json | cut -c1-400
