cd $GRAFT_REPO_ROOT
TAG=${1:-r02k}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python -m tests.parity_report --out gpurun_out/parity_full_$TAG.json > gpurun_out/parity_$TAG.log 2>&1; tail -1 gpurun_out/parity_$TAG.log | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; cut -c1-260 gpurun_out/bench_n1_$TAG.json; tail -2 gpurun_out/bench_n1_$TAG.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
