cd $GRAFT_REPO_ROOT
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown33.log 2>&1
GLASS_DEBUG_SKIP=1 timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown33_skip.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 13 -c 4 -o gpurun_out/conv33_g python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu33g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 66 -c 6 -o gpurun_out/conv33_d python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu33d.log 2>&1
