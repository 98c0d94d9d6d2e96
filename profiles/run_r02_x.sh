# all GPU tests + A/B (previous commit's library vs this tree) + per-layer breakdown
cd $GRAFT_REPO_ROOT
TAG=${1:-r02x}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
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
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "total conv|^G1[1-6]|^D[0167]:|^D:|^C0:" gpurun_out/breakdown_$TAG.log
