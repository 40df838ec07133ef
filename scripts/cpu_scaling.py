import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time
from oracle.oracle import OracleVecEnv, max_threads
print("max_threads", max_threads())
n=65536
env=OracleVecEnv("MiniGrid-DoorKey-8x8-v0", n); env.reset(seed=0)
a=np.random.default_rng(1).integers(0,7,(50,n)).astype(np.int32)
for nt in [1,8,32,64,128]:
    s,_=env.rollout(a, n_threads=nt); print(nt, "threads", n*50/s/1e6, "M steps/s")
