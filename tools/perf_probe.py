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
        st.step(np.float32(1/120.), iters, flags or int(os.environ.get('PROBE_FLAGS', '0')))
        s_ = st.stats()
        if s >= 5:
            v = np.array([s_.msWorld, s_.msBroad, s_.msNarrow, s_.msInit, s_.msSolve, s_.msIntegrate, s_.msTotal])
            acc = v if acc is None else acc + v; n += 1
    acc /= n
    print("%s N=%d pairs=%d manifolds=%d contacts=%d colors=%d cells=%d cellSize=%.3f big=%d launches=%d" % (label, nb, s_.numBroadphaseOverlaps, s_.numCollisions, s_.numContacts, s_.numColors[6], s_.numGridCells, s_.cellSize, s_.numBigColliders, s_.kernelLaunches))
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


def probe_host(config, nx, ny, nz, label, spacing, iters=10, steps=12, mode=0, seed=1578123):
    import time
    from d3d12renderer_b200 import host
    t0 = time.time()
    sc = host.build_scene(config, nx, ny, nz, seed=seed, mode=mode, spacing=spacing)
    print("%s: scene built in %.1f s" % (label, time.time() - t0))
    nb, nc = len(sc["bodies"]), len(sc["colliders"])
    nj = max([len(j[0]) for j in sc["joints"].values()] + [1])
    st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
    if "hull_vertices" in sc:
        st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
    st.upload_bodies(sc["bodies"]); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    for k, (j, p) in sc["joints"].items():
        st.upload_joints(k, j, p)
    st.enable_timing(True)
    acc = None; n = 0
    for s in range(steps):
        st.step(np.float32(1/120.), iters, int(os.environ.get('PROBE_FLAGS', '0')))
        s_ = st.stats(check=False)
        if s >= 3:
            v = np.array([s_.msWorld, s_.msBroad, s_.msNarrow, s_.msInit, s_.msSolve, s_.msIntegrate, s_.msTotal]); acc = v if acc is None else acc + v; n += 1
    acc /= n
    print("%s N=%d C=%d pairs=%d manifolds=%d contacts=%d colors=%s cells=%d cellSize=%.3f big=%d epaOvf=%d gjkCap=%d asserts=%d overflow=(%d,%d,%d)" % (
        label, nb, nc, s_.numBroadphaseOverlaps, s_.numCollisions, s_.numContacts, list(s_.numColors), s_.numGridCells, s_.cellSize, s_.numBigColliders,
        s_.epaOverflows, s_.gjkIterationCaps, s_.assertCount, s_.overflowPairs, s_.overflowManifolds, s_.overflowGrid))
    print("   ms: world %.3f broad %.3f narrow %.3f init %.3f solve %.3f integrate %.3f total %.3f  -> %.1f steps/s" % (*acc, 1000.0/acc[6]))
    bs = st.download_bodies(); print("   finite:", bool(np.all(np.isfinite(bs["position"]))), "y range", float(bs["position"][:,1].min()), float(bs["position"][:,1].max()))
    st.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "big":
    probe_host(3, 64, 32, 64, "cfg3 mixed 131072", 0.8)
    probe_host(3, 128, 64, 128, "cfg3 mixed 1M", 0.8, steps=8)
    probe_host(4, 32, 1, 32, "cfg4 ragdolls 1024", 1.6, steps=40, seed=20260924)
    probe_host(4, 128, 1, 128, "cfg4 ragdolls 16384", 1.6, steps=40, seed=20260924)
