"""EPA kernel timing on the 1 M mixed pile (development aid). Settings come from the environment (B200PHYS_EPA, B200PHYS_EPA_BPS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, host
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sc = host.build_scene(host.CONFIG_MIXED_PILE, n, n // 2, n, seed=1578123, spacing=0.8)
st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=4)
st.set_hulls(sc["hull_vertices"], sc["hull_offsets"]); st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
st.enable_timing(True)
acc = {}
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for s in range(nsteps):
    st.step(np.float32(1 / 120.), 10, 0)
    if s >= nsteps - 5:
        for k, v in st.timeline(): acc[k] = acc.get(k, 0.0) + v / 5
s = st.stats(check=False)
print("EPA=%s BPS=%s  N=%d manifolds=%d epaOvf=%d epaRuns=%d redo=%d redo2=%d | " % (os.environ.get("B200PHYS_EPA", "small"), os.environ.get("B200PHYS_EPA_BPS", "-"), len(sc["bodies"]), s.numCollisions, s.epaOverflows, s.reserved[4], s.reserved[5] & 0xFFFFF, s.reserved[5] >> 20) +
      " ".join("%s %.3f" % (k, v) for k, v in acc.items() if k.startswith("np_")) + " | total %.3f" % sum(acc.values()))
