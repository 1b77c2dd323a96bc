"""Step time with and without CUDA graph replay (B200PHYS_GRAPH=0 disables it): wall clock around N untimed steps + one synchronise."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, scenes, host


def probe(sc, label, steps=200, flags=0):
    nj = max([len(j[0]) for j in sc.get("joints", {}).values()] + [1])
    st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
    if "hull_vertices" in sc:
        st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
    st.upload_bodies(np.ascontiguousarray(sc["bodies"])); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    for k, (j, p) in sc.get("joints", {}).items():
        st.upload_joints(k, j, p)
    for _ in range(40):
        st.step(np.float32(1 / 120.), 10, flags)
    st.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(np.float32(1 / 120.), 10, flags)
    st.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    s = st.stats(check=False)
    print("%-28s graph=%s  %.4f ms/step  launches/step %d  replays %d captures %d contacts %d" % (label, os.environ.get("B200PHYS_GRAPH", "1"), ms, s.kernelLaunches, s.reserved[6], s.reserved[0], s.numContacts))
    st.close()


if __name__ == "__main__":
    probe(scenes.spheres_on_ground(32, mode="rest"), "cfg1 1024 spheres", 400)
    probe(host.build_scene(4, 8, 1, 8, seed=20260924, mode=0, spacing=1.6), "64 ragdolls", 400)
    probe(scenes.sphere_pile_settled(64, 64, 64), "cfg2 262144 spheres", 60)
    probe(host.build_scene(4, 128, 1, 128, seed=20260924, mode=0, spacing=1.6), "cfg4 16384 ragdolls", 60)
