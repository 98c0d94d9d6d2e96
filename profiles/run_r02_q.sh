cd $GRAFT_REPO_ROOT
TAG=${1:-r02q}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tests/profile_step.py --pop 64 --evals 41 2>&1 | grep "step ms" > gpurun_out/step_$TAG.log
python - <<PY
import re,statistics
v=[float(t) for t in re.findall(r"\d+\.\d+", open('gpurun_out/step_$TAG.log').read().split("eval:")[1])][1:]
print("n",len(v),"median",round(statistics.median(v),2),"mean",round(sum(v)/len(v),2))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fir --csv --log-file gpurun_out/fir_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/fir_$TAG.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
per={}
for r in rows[h+1:]:
    per.setdefault(r[0],{'name':r[4][:36]})[r[12]]=float(r[14].replace(',',''))
tot=0
for k,d in per.items():
    t=d['gpu__time_duration.sum']/1e3; tot+=t
    print(d['name'], round(t,1),'us issue', d['smsp__issue_active.avg.pct_of_peak_sustained_active'])
print("total", round(tot))
PY
