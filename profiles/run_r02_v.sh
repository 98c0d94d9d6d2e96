# ncu --set full (with source) of three hot kernels: last generator conv, G15 folded up-conv, fused D down-conv 64->128
cd $GRAFT_REPO_ROOT
TAG=${1:-r02v}
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:conv_tc_kernel<\(int\)32, \(int\)32, \(int\)4, \(int\)3>|conv_tc_kernel<\(int\)64, \(int\)64, \(int\)4, \(int\)2>|downconv_tc_kernel<\(int\)64, \(int\)128>' -c 3 \
  -o gpurun_out/hot3_$TAG python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -3 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/hot3_$TAG.ncu-rep
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "total conv|^G1[1-6]|^D[01]:" gpurun_out/breakdown_$TAG.log
