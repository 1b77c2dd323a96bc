"""Throughput of the batched locomotion environment (host/locomotion_env.cpp): environment steps per second."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import host
for n in (int(a) for a in (sys.argv[1:] or ["1024", "4096"])):
    env = host.LocomotionEnv(n, spacing=4.0)
    env.reset()
    a = np.zeros((n, 27), dtype=np.float32)
    for _ in range(5):
        env.step(a)
    t = time.time(); k = 20
    for _ in range(k):
        s, r, f = env.step(a)
    dt = (time.time() - t) / k
    print("envs %d: %.2f ms per batched step -> %.0f environment steps/s (failed %d)" % (n, dt * 1e3, n / dt, int(f.sum())))
    env.close()
