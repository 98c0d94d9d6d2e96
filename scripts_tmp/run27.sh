cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"from_rgb_fir|fir_down|blur_s2d|rgb_combine|upfir|attention" -c 14 -o gpurun_out/hbm27 python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu27.log 2>&1
ls -la gpurun_out/
