"""Stress / determinism record for the barrier-free PGS sweep (k_solver.cu: per-body update counters, relaxed 16-byte loads / stores).
The protocol has no fences and its correctness argument rests on the two 16-byte halves of a body record being self-validating; the
memory model does not promise that, the hardware does. Evidence kept under profiles/: the same scene stepped REPS times from the same
state, in the fast order and in the exact order, must end in bit-identical body states every time (any torn or stale read of a velocity
record would change low-order bits of the result), and the exact order must equal a second context's result as well."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from d3d12renderer_b200 import abi, host, scenes

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
STEPS = 8


def run(sc, flags):
    nj = max([len(j[0]) for j in sc.get("joints", {}).values()] + [1])
    st = abi.Stepper(max_bodies=len(sc["bodies"]), max_colliders=len(sc["colliders"]), max_transforms=max(4, len(sc["transforms"])), max_joints=nj)
    if "hull_vertices" in sc:
        st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
    st.upload_bodies(np.ascontiguousarray(sc["bodies"])); st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"])
    for k, (j, p) in sc.get("joints", {}).items():
        st.upload_joints(k, j, p)
    for _ in range(STEPS):
        st.step(np.float32(1 / 120.), 10, flags)
    out = st.download_bodies()
    s = st.stats(check=False)
    st.close()
    return hashlib.sha256(out.tobytes()).hexdigest()[:16], int(s.numContacts)


cases = [("262144-sphere pile", lambda: scenes.sphere_pile_settled(64, 64, 64)),
         ("131072 mixed primitives", lambda: host.build_scene(3, 64, 32, 64, seed=1578123, mode=0, spacing=0.8)),
         ("4096 ragdolls", lambda: host.build_scene(4, 64, 1, 64, seed=20260924, mode=0, spacing=1.6))]
bad = 0
for name, make in cases:
    sc = make()
    for flags, label in ((abi.STEP_FAST, "fast"), (abi.STEP_EXACT_ORDER, "exact")):
        reps = REPS if flags == abi.STEP_FAST else max(3, REPS // 8)
        digests = set()
        for r in range(reps):
            d, k = run(sc, flags)
            digests.add(d)
        print("%-26s %-5s order: %3d runs x %d steps, %7d contacts, distinct results: %d  %s" % (name, label, reps, STEPS, k, len(digests), sorted(digests)[0]))
        bad += len(digests) != 1
print("RESULT:", "every run bit-identical" if not bad else "NON-DETERMINISTIC RESULTS")
sys.exit(1 if bad else 0)
