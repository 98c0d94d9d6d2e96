cd $GRAFT_REPO_ROOT
for v in "" _p3 _p4; do
  export CLIPGLASS_LIB=$PWD/clip_glass_b200/libclipglass_b200$v.so
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"upfir|blur_s2d" --csv --log-file gpurun_out/poly47$v.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
  echo "== $v"; grep -E "upfir|blur" gpurun_out/poly47$v.csv | awk -F'","' '{print substr($5,1,30), $NF}'
done
