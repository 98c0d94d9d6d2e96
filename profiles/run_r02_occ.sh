# A/B: resident blocks per SM of the streaming FIR passes (default build vs GLASS_FIR_MINB=4 GLASS_BLUR_MINB=4 GLASS_POLY_MINB=2)
cd $GRAFT_REPO_ROOT
for lib in libclipglass_b200.so libclipglass_b200_occ4.so; do
  CLIPGLASS_LIB=clip_glass_b200/$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fir|blur' --csv python tests/profile_step.py --pop 64 --evals 1 2>/dev/null | grep -E "fir|blur" | awk -F'","' -v l=$lib '{gsub(/\(.*/,"",$5); gsub(/"/,"",$NF); printf "%s %s %s\n", l, $5, $NF}'
done | tee gpurun_out/occ_ab.log
CLIPGLASS_LIB=clip_glass_b200/libclipglass_b200_occ4.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for round in 1 2; do
  timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/variant A (this tree) /"
  CLIPGLASS_LIB=clip_glass_b200/libclipglass_b200_occ4.so timeout 300 python tests/profile_step.py --pop 64 --evals 21 2>&1 | grep "step ms" | sed "s/^/variant B (occ4)      /"
done > gpurun_out/ab_occ4.log
python - <<PY
import re,statistics,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/ab_occ4.log'):
    d[l[:21].strip()]+=[float(t) for t in re.findall(r"\d+\.\d+", l.split("eval:")[1])][1:]
for k,v in d.items(): print(k, "n",len(v),"median",round(statistics.median(v),2),"mean",round(sum(v)/len(v),2))
PY
