cd $GRAFT_REPO_ROOT
TAG=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; cut -c1-260 gpurun_out/bench_n1_$TAG.json
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "step ms|total conv" gpurun_out/breakdown_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/conv_dram_$TAG.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
wc -l gpurun_out/launches_$TAG.csv gpurun_out/conv_dram_$TAG.csv
