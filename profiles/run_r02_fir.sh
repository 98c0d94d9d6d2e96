# parity tests + ncu per-launch time / issue utilisation of the stride-2 FIR passes (fromRGB+FIR, projection FIRs, upfir, blur)
cd $GRAFT_REPO_ROOT
TAG=${1:-r02fir}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'fir|blur|rgb_combine|resize|noise|vecmat|layernorm|attention|embed|mbstd|dense1|cosine|pixelnorm|style|const_input' --csv --log-file gpurun_out/fir_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/fir_$TAG.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
per=collections.OrderedDict()
for r in rows[h+1:]:
    per.setdefault(r[0],{'name':r[4].replace('void ','').replace('unnamed>::','')[:34]})[r[12]]=float(r[14].replace(',',''))
agg=collections.OrderedDict()
for k,d in per.items():
    a=agg.setdefault(d['name'],[0,0.0,0.0,0.0]); a[0]+=1; a[1]+=d['gpu__time_duration.sum']/1e3
    a[2]+=d.get('dram__bytes_read.sum',0)+d.get('dram__bytes_write.sum',0); a[3]=max(a[3],d['smsp__issue_active.avg.pct_of_peak_sustained_active'])
tot=0
for n,a in agg.items():
    tot+=a[1]; print(f"{n:36s} x{a[0]:3d} {a[1]:8.1f} us  {a[2]/1e6 if a[2]<1e12 else a[2]/1e6:10.1f} MB(or unit)  issue<= {a[3]:.0f}%")
print("total non-conv us", round(tot))
PY
