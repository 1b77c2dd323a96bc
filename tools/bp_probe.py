"""Broad-phase timing (development aid): kernel-group timeline for the sphere pile, the mixed pile and the ragdoll field. Env: B200PHYS_CELL_DIV."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, host, scenes
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def run(name, sc):
    nj = max([len(j[0]) for j in sc.get("joints", {}).values()] + [1])
    st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
    if "hull_vertices" in sc: st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
    st.upload_bodies(np.ascontiguousarray(sc["bodies"])); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    for k, (j, p) in sc.get("joints", {}).items(): st.upload_joints(k, j, p)
    st.enable_timing(True)
    acc = {}
    for s in range(10):
        st.step(np.float32(1 / 120.), 10, 0)
        if s >= 4:
            for k, v in st.timeline(): acc[k] = acc.get(k, 0.0) + v / 6
    s = st.stats(check=False)
    print("%-14s div=%s C=%d pairs=%d cells=%d cell=%.3f big=%d | " % (name, os.environ.get("B200PHYS_CELL_DIV", "auto"), len(sc["colliders"]), s.numBroadphaseOverlaps, s.numGridCells, s.cellSize, s.numBigColliders) +
          " ".join("%s %.3f" % (k, v) for k, v in acc.items() if k.startswith("bp_") or k == "world_colliders") + " | step %.3f launches %d" % (sum(acc.values()), s.kernelLaunches))
    st.close()
if which in ("all", "cfg2"): run("spheres 262k", scenes.sphere_pile_settled(64, 64, 64, seed=519431))
if which in ("all", "cfg3"): run("mixed 1M", host.build_scene(host.CONFIG_MIXED_PILE, 128, 64, 128, seed=1578123, spacing=0.8))
if which in ("all", "cfg4"): run("ragdolls 16k", host.build_scene(host.CONFIG_RAGDOLLS, 128, 1, 128, seed=20260924, mode=0, spacing=1.6))
if which in ("all", "cfg1"): run("spheres 1024", scenes.spheres_on_ground(32, mode="rest"))
