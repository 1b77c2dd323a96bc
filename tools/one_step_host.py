"""Runs a few steps of a host-built config (3 = mixed primitives, 4 = ragdolls) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, host
config = int(sys.argv[1]); nx, ny, nz = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]); steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
sc = host.build_scene(config, nx, ny, nz, seed=1578123, mode=0, spacing=0.8 if config == 3 else 1.6)
nj = max([len(j[0]) for j in sc["joints"].values()] + [1])
st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
if "hull_vertices" in sc:
    st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
for k, (j, p) in sc["joints"].items():
    st.upload_joints(k, j, p)
for _ in range(steps):
    st.step(np.float32(1 / 120.), 10, 0)
st.synchronize()
s = st.stats(check=False); print("contacts", s.numContacts, "colors", list(s.numColors), "launches", s.kernelLaunches)
