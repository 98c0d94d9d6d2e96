"""How many host threads make the CPU oracle fastest on this box? (bench.py's cpu_baseline uses the best)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for t in (8, 16, 32, 64, os.cpu_count()):
    bench.cpu_oracle_step(4, 4, True, 1, t)
    dt = bench.cpu_oracle_step(4, 4, True, 2, t)
    print(f"threads={t}: {dt:.2f} s per 4 candidates -> {4/dt:.3f} cand/s", flush=True)
