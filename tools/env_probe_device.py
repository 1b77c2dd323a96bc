"""Environment steps per second of the batched locomotion environment: host-side formulas vs everything on the device (k_env.cu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import host
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for device in (True, False):
    env = host.LocomotionEnv(E, spacing=6.0)
    env.set_device_mode(device)
    env.reset()
    lo, hi = env.action_ranges()
    rng = np.random.default_rng(1)
    a = (0.2 * rng.uniform(lo, hi, size=(E, len(lo)))).astype(np.float32)
    for _ in range(5):
        env.step(a)
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        s, r, f = env.step(a)
    dt = (time.perf_counter() - t0) / n
    print("%d environments, %s: %.2f ms per step = %.2f M environment steps/s (failed %d)" % (E, "device mode" if device else "host mode", dt * 1e3, E / dt / 1e6, int(f.sum())))
    env.close()
