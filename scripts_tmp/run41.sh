cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 15 -c 2 -o gpurun_out/conv41_g python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu41g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 66 -c 3 -o gpurun_out/conv41_d python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu41d.log 2>&1
