cd $GRAFT_REPO_ROOT
TAG=${1:-r02n}
for fl in 0 2048; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:blur_s2d --csv --log-file gpurun_out/blur_${TAG}_$fl.csv python tests/profile_step.py --pop 64 --evals 1 --flags $fl > /dev/null 2>&1
done
python - <<PY
import csv
for fl in (0,2048):
    rows=list(csv.reader(open(f'gpurun_out/blur_${TAG}_{fl}.csv')))
    h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
    per={}
    for r in rows[h+1:]:
        per.setdefault(r[0],{'name':r[4][:40]})[r[12]]=float(r[14].replace(',',''))
    for k,d in per.items():
        t=d['gpu__time_duration.sum']/1e3; b=d['dram__bytes_read.sum']+d['dram__bytes_write.sum']
        print(fl, d['name'], round(t,1),'us', round(b/1e6),'MB', round(b/t/1e3),'GB/s issue', d['smsp__issue_active.avg.pct_of_peak_sustained_active'])
PY
