# projection as a second accumulator of the fused down-conv (default) vs the separate 1x1 GEMM (flags 4096): parity + A/B
cd $GRAFT_REPO_ROOT
TAG=${1:-r02proj}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
for round in 1 2 3; do
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 --flags 4096 2>&1 | grep "step ms" | sed "s/^/variant A (separate GEMM) /"
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/variant B (accumulator)  /"
done > gpurun_out/ab_$TAG.log
python - <<PY
import re,statistics,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/ab_$TAG.log'):
    d[l[:25].strip()]+=[float(t) for t in re.findall(r"\d+\.\d+", l.split("eval:")[1])][1:]
for k,v in d.items(): print(k, "n",len(v),"median",round(statistics.median(v),2),"mean",round(sum(v)/len(v),2))
PY
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing --flags 4096 2>&1 | grep -E "^D0:|total conv" | sed "s/^/A /"
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing 2>&1 | grep -E "^D0:|total conv" | sed "s/^/B /"
