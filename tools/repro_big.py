import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, host
nx, ny, nz = [int(x) for x in sys.argv[1:4]]
sc = host.build_scene(3, nx, ny, nz, seed=1578123, spacing=0.8)
nb, nc = len(sc["bodies"]), len(sc["colliders"])
st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=4)
st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
for s in range(int(sys.argv[4]) if len(sys.argv) > 4 else 1):
    st.step(np.float32(1/120.), int(os.environ.get("ITERS", "2")), 0); st.synchronize()
    x = st.stats(check=False); print("step", s, "pairs", x.numBroadphaseOverlaps, "manifolds", x.numCollisions, "contacts", x.numContacts, "overflow", x.overflowPairs, x.overflowManifolds, x.overflowGrid, flush=True)
