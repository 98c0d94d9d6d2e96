cd $GRAFT_REPO_ROOT
timeout 300 python -m tests.gpu_diag --config tiny --impl 0 > gpurun_out/diag28_tc.log 2>&1; grep -E "block|features|sim_vs|neg_sim|hinge" gpurun_out/diag28_tc.log | cut -c1-200
timeout 300 python -m tests.gpu_diag --config tiny --impl 0 --flags 16 > gpurun_out/diag28_simt.log 2>&1; grep -E "block|features|sim_vs" gpurun_out/diag28_simt.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu28.log 2>&1; tail -5 gpurun_out/pytest_gpu28.log
timeout 300 python tests/profile_step.py --pop 64 --evals 5 2>&1 | grep "step ms"
timeout 300 python tests/profile_step.py --pop 64 --evals 5 --flags 16 2>&1 | grep "step ms"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"from_rgb_fir|fir_down|blur_s2d|attention_tc" -c 9 -o gpurun_out/hbm28 python tests/profile_step.py --pop 64 --evals 1 > gpurun_out/ncu28.log 2>&1
