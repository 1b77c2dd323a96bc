import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from d3d12renderer_b200 import host
import ref_lib
ref = ref_lib.RefEnv(); env = host.LocomotionEnv(4, spacing=6.0); env.set_exact_order(True); env.set_push_probability(0.0)
lo, hi = env.action_ranges()
for scale in (0.0, 0.05):
    rng = np.random.default_rng(12)
    env.reset(); ref.reset(seed=77, push_probability=0.0, simd=False)
    print("scale", scale)
    for step in range(30):
        a = (scale * rng.uniform(lo, hi)).astype(np.float32)
        s, r, f = env.step(np.tile(a, (4, 1))); rs, rr, rf = ref.step(a)
        d = np.abs(s[0] - rs); k = int(np.argmax(d))
        print(" step %2d maxerr %.2e at %2d (ours %.5f ref %.5f) reward %.5f / %.5f  head y %.4f contacts? toes y %.4f" % (step, d.max(), k, s[0][k], rs[k], r[0], rr, rs[22], rs[4]))
