# round 2, call A: full GPU test suite, parity table, bench at N=1, per-layer breakdown
cd $GRAFT_REPO_ROOT
TAG=${1:-r02a}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python -m tests.parity_report --out gpurun_out/parity_full_$TAG.json > gpurun_out/parity_$TAG.log 2>&1; tail -3 gpurun_out/parity_$TAG.log | cut -c1-400
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; cut -c1-300 gpurun_out/bench_n1_$TAG.json; tail -2 gpurun_out/bench_n1_$TAG.err
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown_$TAG.log 2>&1; grep -E "step ms|total conv" gpurun_out/breakdown_$TAG.log
