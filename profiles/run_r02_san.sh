# compute-sanitizer memcheck + racecheck: every CUDA path at the tiny configuration, and one eager full-size evaluation
cd $GRAFT_REPO_ROOT
TAG=${1:-r02san}
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_step.py > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck tiny rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_step.py --full > gpurun_out/sanitizer_memcheck_full_$TAG.log 2>&1; echo "memcheck full rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_full_$TAG.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_step.py > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck tiny rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_$TAG.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_step.py --full > gpurun_out/sanitizer_racecheck_full_$TAG.log 2>&1; echo "racecheck full rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_full_$TAG.log
