cd $GRAFT_REPO_ROOT
TAG=${1:-r02l}
FLAGB=${2:-1024}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact or multi_seed or resampling or benchmarked" > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
for round in 1 2 3; do
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/variant A (default) /"
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 --flags $FLAGB 2>&1 | grep "step ms" | sed "s/^/variant B (flags)   /"
done > gpurun_out/ab_$TAG.log
python - <<PY
import re,statistics,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/ab_$TAG.log'):
    d[l[:19].strip()]+=[float(t) for t in re.findall(r"\d+\.\d+", l.split("eval:")[1])][1:]
for k,v in d.items(): print(k, "n",len(v),"median",round(statistics.median(v),2),"mean",round(sum(v)/len(v),2))
PY
