"""Runs a few steps of the bench workload (for ncu captures: keep it short)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenes.sphere_pile_settled(n, n, n, seed=519431)
st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=4)
st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
for _ in range(steps):
    st.step(np.float32(1/120.), 10, 0)
st.synchronize()
s = st.stats(); print("contacts", s.numContacts, "colors", list(s.numColors), "launches", s.kernelLaunches)
