# gpurun recipe of the GPU-resident GA evidence (one B200): the glass_ga_* kernel tests, the search-loop measurement,
# a bench line of the same build and the rest of the GPU suite.  Most important first: the call may be cut short.
cd $GRAFT_REPO_ROOT
timeout 140 python -m pytest tests/test_gpu_ga.py -m gpu -q --tb=short --durations=3 > gpurun_out/r02_pytest_gpu_ga.log 2>&1; tail -25 gpurun_out/r02_pytest_gpu_ga.log
timeout 110 python tests/profile_ga.py --gens 20 > gpurun_out/r02_ga_loop_p64.json 2> gpurun_out/r02_ga_loop_p64.err; cat gpurun_out/r02_ga_loop_p64.json; tail -3 gpurun_out/r02_ga_loop_p64.err
timeout 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_ga_build.json 2> gpurun_out/r02_bench_n1_ga_build.err; cut -c1-400 gpurun_out/r02_bench_n1_ga_build.json
timeout 200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_ga.py > gpurun_out/r02_pytest_gpu_rest.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_rest.log
