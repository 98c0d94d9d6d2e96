cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu31.log 2>&1; tail -5 gpurun_out/pytest_gpu31.log
echo eager; GLASS_DEBUG_NO_GRAPH=1 timeout 300 python tests/profile_step.py --pop 64 --evals 10 2>&1 | grep "step ms"
echo graph-nofork; GLASS_DEBUG_NO_FORK=1 timeout 300 python tests/profile_step.py --pop 64 --evals 10 2>&1 | grep "step ms"
echo graph-fork; timeout 300 python tests/profile_step.py --pop 64 --evals 10 2>&1 | grep "step ms"
echo eager; GLASS_DEBUG_NO_GRAPH=1 timeout 300 python tests/profile_step.py --pop 64 --evals 10 2>&1 | grep "step ms"
echo graph-fork; timeout 300 python tests/profile_step.py --pop 64 --evals 10 2>&1 | grep "step ms"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench31.json 2> gpurun_out/bench31.err; cut -c1-300 gpurun_out/bench31.json
