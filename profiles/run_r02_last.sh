# gpurun recipe: launch list (ncu, one evaluation at P = 64) and a full bench line of the final build
cd $GRAFT_REPO_ROOT
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_p64_last.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1; wc -l gpurun_out/r02_launches_p64_last.csv
timeout 75 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_last.json 2> gpurun_out/r02_bench_n1_last.err; cut -c1-200 gpurun_out/r02_bench_n1_last.json
timeout 40 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/r02_conv_breakdown_p64_last.log 2>&1; grep -E "step ms|total conv" gpurun_out/r02_conv_breakdown_p64_last.log
