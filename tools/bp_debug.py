import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from d3d12renderer_b200 import abi, scenes
import oracle_lib
for name, sc in [("kat", scenes.spheres_on_ground(1, mode="rest")), ("4x4", scenes.spheres_on_ground(4, mode="rest")), ("pile", scenes.sphere_pile_settled(6,4,6))]:
    st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=4)
    st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    st.step(np.float32(1/120), 10, 0)
    s = st.stats(check=False)
    p = st.download_pairs()
    o = oracle_lib.Oracle(sc); o.step(np.float32(1/120), 10)
    print(name, "gpu pairs", len(p), "oracle", len(o.pairs()), "equal", np.array_equal(p, o.pairs()), "cells", s.numGridCells, "cellSize", s.cellSize, "big", s.numBigColliders)
    if not np.array_equal(p, o.pairs()): print(p[:10], o.pairs()[:10])
    st.close()
