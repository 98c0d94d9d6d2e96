cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu26.log 2>&1; tail -5 gpurun_out/pytest_gpu26.log
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --timing > gpurun_out/breakdown26.log 2>&1; head -3 gpurun_out/breakdown26.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches26.csv python tests/profile_step.py --pop 64 --evals 1 > /dev/null 2>&1
wc -l gpurun_out/launches26.csv
python bench.py --steps 10 --warmup 3 > gpurun_out/bench26.json 2> gpurun_out/bench26.err; cut -c1-300 gpurun_out/bench26.json
