import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from d3d12renderer_b200 import abi, host
sc = host.build_scene(4, 128, 1, 128, seed=20260924, mode=0, spacing=1.6)
nj = max([len(j[0]) for j in sc["joints"].values()] + [1])
st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
for k, (j, p) in sc["joints"].items(): st.upload_joints(k, j, p)
st.enable_timing(True)
for s in range(30): st.step(np.float32(1/120.), 10, 0)
print(dict(st.timeline()))
