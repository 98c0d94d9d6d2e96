# round 2 final evidence for the committed build: GPU tests, bench lines (both workloads), per-layer breakdown,
# ncu launch list + per-launch metrics of one step
cd $GRAFT_REPO_ROOT
TAG=${1:-r02final}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; cut -c1-250 gpurun_out/bench_n1_$TAG.json; tail -1 gpurun_out/bench_n1_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_reference_$TAG.json
python bench.py --workload gpt2 --steps 10 --warmup 3 > gpurun_out/bench_gpt2_$TAG.json 2> gpurun_out/bench_gpt2_$TAG.err; cut -c1-250 gpurun_out/bench_gpt2_$TAG.json
timeout 300 python tests/profile_step.py --pop 64 --evals 4 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "total conv" gpurun_out/breakdown_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/metrics_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
wc -l gpurun_out/launches_$TAG.csv gpurun_out/metrics_$TAG.csv
