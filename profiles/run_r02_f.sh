# round 2, call F: evidence — ncu launch list, DRAM traffic of the conv launches, tensor-pipe / DRAM metrics per launch,
# compute-sanitizer memcheck + racecheck of the tiny configuration
cd $GRAFT_REPO_ROOT
TAG=${1:-r02f}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/metrics_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
wc -l gpurun_out/launches_$TAG.csv gpurun_out/metrics_$TAG.csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_step.py > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_step.py > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_$TAG.log
