cd $GRAFT_REPO_ROOT
TAG=${1:-r02p}
timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -x -q > gpurun_out/pytest_text_$TAG.log 2>&1; tail -5 gpurun_out/pytest_text_$TAG.log
python tests/profile_text.py --pop 64 --evals 5 2>&1 | tail -3
