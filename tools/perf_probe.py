"""Quick per-stage timing probe (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, scenes

def probe(sc, iters=10, steps=30, label="", flags=0):
    nb, nc = len(sc["bodies"]), len(sc["colliders"])
    st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=4)
    st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    st.enable_timing(True)
    acc = None; n = 0
    for s in range(steps):
        st.step(np.float32(1/120.), iters, flags)
        s_ = st.stats()
        if s >= 5:
            v = np.array([s_.msWorld, s_.msBroad, s_.msNarrow, s_.msInit, s_.msSolve, s_.msIntegrate, s_.msTotal])
            acc = v if acc is None else acc + v; n += 1
    acc /= n
    print("%s N=%d pairs=%d manifolds=%d contacts=%d colors=%d cells=%d cellSize=%.3f big=%d launches=%d rounds=%d" % (label, nb, s_.numBroadphaseOverlaps, s_.numCollisions, s_.numContacts, s_.numColors[6], s_.numGridCells, s_.cellSize, s_.numBigColliders, s_.kernelLaunches, s_.reserved[0]))
    print("   ms: world %.3f broad %.3f narrow %.3f init %.3f solve %.3f integrate %.3f total %.3f  -> %.1f steps/s" % (*acc, 1000.0/acc[6]))
    st.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "cfg1"):
        probe(scenes.spheres_on_ground(32, mode="rest"), label="cfg1 rest")
    if which in ("all", "cfg2"):
        probe(scenes.sphere_box_drop(64, 64, 64), label="cfg2 freefall 262144")
        probe(scenes.sphere_pile_settled(64, 64, 64), label="cfg2 dense 262144", steps=12)
    if which in ("all", "1m"):
        probe(scenes.sphere_pile_settled(128, 64, 128), label="dense 1M", steps=10)
