cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"from_rgb_fir|fir_down|blur_s2d" -c 12 -o gpurun_out/hbm30 python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu30.log 2>&1
