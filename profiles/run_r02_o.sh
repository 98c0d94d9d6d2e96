cd $GRAFT_REPO_ROOT
TAG=${1:-r02o}
python tests/profile_text.py --pop 64 --evals 5 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/text_launches_$TAG.csv python tests/profile_text.py --pop 64 --evals 1 > /dev/null 2>&1
python - <<PY
import csv, collections, re
rows=list(csv.reader(open('gpurun_out/text_launches_$TAG.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
agg=collections.defaultdict(lambda:[0,0.0])
per=[]
for r in rows[h+1:]:
    name=re.sub(r"^void ","",r[4]).split("(")[0].split("::")[-1]
    t=float(r[-1].replace(",",""))/1e3
    agg[name+" grid"+r[8]][0]+=1; agg[name+" grid"+r[8]][1]+=t
tot=sum(v[1] for v in agg.values())
print("total us", round(tot))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]: print(f"{v[1]:9.0f} us {v[0]:5d} x {v[1]/v[0]:7.1f} us  {k}")
PY
