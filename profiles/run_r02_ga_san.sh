# gpurun recipe: compute-sanitizer memcheck over the glass_ga_* kernel tests (the run.py end-to-end tests, which are
# dominated by the fitness engine already covered by profiles/r02_sanitizer_*.log, are left out)
cd $GRAFT_REPO_ROOT
timeout 95 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ga.py -m gpu -q -x -k "not run_driver" > gpurun_out/r02_sanitizer_memcheck_ga.log 2>&1; tail -6 gpurun_out/r02_sanitizer_memcheck_ga.log
# second call: racecheck (shared-memory hazards: permutation keys, the counters of the duplicate / survival kernels)
timeout 45 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ga.py -m gpu -q -x -k "not run_driver" > gpurun_out/r02_sanitizer_racecheck_ga.log 2>&1
