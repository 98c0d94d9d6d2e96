# parity tests + ncu --set full (with source) of three hot kernels: last generator conv, G15 folded up-conv, fused D down-conv 64->128
cd $GRAFT_REPO_ROOT
TAG=${1:-r02u}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:conv_tc_kernel<32, 32, 4, 3>|conv_tc_kernel<64, 64, 4, 2>|downconv_tc_kernel<64, 128>' -c 3 \
  -o gpurun_out/hot3_$TAG python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -3 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/hot3_$TAG.ncu-rep
for round in 1 2; do
  CLIPGLASS_LIB=clip_glass_b200/libclipglass_b200_prev.so timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/variant A (previous) /"
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/variant B (this tree)/"
done > gpurun_out/ab_$TAG.log
python - <<PY
import re,statistics,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/ab_$TAG.log'):
    d[l[:21].strip()]+=[float(t) for t in re.findall(r"\d+\.\d+", l.split("eval:")[1])][1:]
for k,v in d.items(): print(k, "n",len(v),"median",round(statistics.median(v),2),"mean",round(sum(v)/len(v),2))
PY
