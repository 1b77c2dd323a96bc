"""Small run of every device path for compute-sanitizer (memcheck): spheres, mixed primitives, ragdolls, events, triggers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, scenes, host

def run(sc, steps, flags=0, events=False, ff=False):
    nj = max([len(j[0]) for j in sc.get("joints", {}).values()] + [1])
    st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
    if "hull_vertices" in sc:
        st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
    st.upload_bodies(np.ascontiguousarray(sc["bodies"])); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    for k, (j, p) in sc.get("joints", {}).items():
        st.upload_joints(k, j, p)
    if events:
        st.enable_events(True)
    if ff:
        st.upload_force_fields(sc["force_fields"], sc["global_force"])
    for s in range(steps):
        st.step(np.float32(1 / 120.), 8, flags)
        if events:
            st.download_events()
        if ff:
            st.download_trigger_events()
    s_ = st.stats(check=False)
    print(sc.get("name", "?"), "flags", flags, "contacts", s_.numContacts, "ok")
    st.close()

run(scenes.sphere_pile_settled(12, 8, 12), 12, 0, events=True)
run(scenes.sphere_pile_settled(12, 8, 12), 3, 1)
run(scenes.sphere_pile_settled(12, 8, 12), 3, 2)
run(host.build_scene(3, 8, 6, 8, seed=1578123, mode=0, spacing=0.8), 10, 0)
run(host.build_scene(4, 6, 1, 6, seed=20260924, mode=0, spacing=1.6), 10, 0)
run(scenes.event_zoo(), 30, 0, ff=True)
