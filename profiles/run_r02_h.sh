cd $GRAFT_REPO_ROOT
TAG=${1:-r02h}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:downconv -c 2 -o gpurun_out/downconv_$TAG python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -3 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/downconv_$TAG.ncu-rep
