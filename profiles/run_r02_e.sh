cd $GRAFT_REPO_ROOT
TAG=${1:-r02e}
timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -x -q > gpurun_out/pytest_text_$TAG.log 2>&1; tail -5 gpurun_out/pytest_text_$TAG.log
python bench.py --workload gpt2 --steps 10 --warmup 3 > gpurun_out/bench_gpt2_$TAG.json 2> gpurun_out/bench_gpt2_$TAG.err; cut -c1-330 gpurun_out/bench_gpt2_$TAG.json; tail -3 gpurun_out/bench_gpt2_$TAG.err
