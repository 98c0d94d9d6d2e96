cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_r02.json 2> gpurun_out/bench_n${N}_r02.err; cut -c1-300 gpurun_out/bench_n${N}_r02.json; tail -3 gpurun_out/bench_n${N}_r02.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload gpt2 --steps 5 --warmup 3 > gpurun_out/bench_gpt2_n${N}_r02.json 2> gpurun_out/bench_gpt2_n${N}_r02.err; cut -c1-300 gpurun_out/bench_gpt2_n${N}_r02.json; tail -3 gpurun_out/bench_gpt2_n${N}_r02.err
