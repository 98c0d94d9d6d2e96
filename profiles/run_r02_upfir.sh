# A/B: k_upfir at one (default) vs two resident blocks per SM (GLASS_POLY_MINB=2 build), ncu time per launch
cd $GRAFT_REPO_ROOT
for lib in libclipglass_b200.so libclipglass_b200_minb2.so; do
  CLIPGLASS_LIB=clip_glass_b200/$lib timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:upfir --csv python tests/profile_step.py --pop 64 --evals 2 2>/dev/null | grep upfir | awk -F'","' -v l=$lib '{print l, $13, $NF}'
done | tee gpurun_out/upfir_ab.log
