"""Per-step solve-kernel time of the dense 262144-sphere pile (development aid: run-to-run / step-to-step spread)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, scenes
sc = scenes.sphere_pile_settled(64, 64, 64)
st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=4)
st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
st.enable_timing(True)
t = []
for s in range(24):
    st.step(np.float32(1 / 120.), 10, 0)
    t.append(st.stats().msSolve)
print("solve ms per step:", " ".join("%.3f" % x for x in t))
