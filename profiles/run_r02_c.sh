cd $GRAFT_REPO_ROOT
TAG=${1:-r02c}
timeout 900 python -m pytest tests/test_gpu_text.py -m gpu -x -q > gpurun_out/pytest_text_$TAG.log 2>&1; tail -25 gpurun_out/pytest_text_$TAG.log
