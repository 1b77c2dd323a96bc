"""Generates the frozen fixtures under tests/golden/ from the CPU oracle (run once, outputs committed).

The reference has no tests or golden vectors and cannot be compiled in this environment, so these files freeze
the ORACLE's outputs; the oracle itself is pinned by the hand-derived known answers in tests/test_cpu_oracle.py
(SURVEY.md Appendix C). A fixture holds the scene inputs (boundary arrays) and, after `steps` steps with canonical pair order
and array-order Gauss-Seidel, the pair list, contacts and body states.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from d3d12renderer_b200 import host, scenes  # noqa: E402
import oracle_lib  # noqa: E402

DT = np.float32(1.0 / 120.0)

CASES = {
    "config1_rest_64": (lambda: scenes.spheres_on_ground(8, mode="rest"), 1, 10),
    "config1_drop_64": (lambda: scenes.spheres_on_ground(8, mode="drop"), 3, 10),
    "sphere_pile_5x4x5": (lambda: scenes.sphere_pile_settled(5, 4, 5), 2, 10),
    "mixed_pile_4x3x4": (lambda: host.build_scene(host.CONFIG_MIXED_PILE, 4, 3, 4, seed=1578123, spacing=0.75), 2, 10),
    "ragdolls_2": (lambda: host.build_scene(host.CONFIG_RAGDOLLS, 2, 1, 1, seed=20260924, mode=0, spacing=2.5), 40, 10),
}


def pack_scene(sc):
    out = {"bodies": sc["bodies"], "colliders": sc["colliders"], "transforms": sc["transforms"]}
    for k, (j, p) in sc.get("joints", {}).items():
        out["joints%d" % k] = j; out["joint_pairs%d" % k] = p
    if "hull_vertices" in sc:
        out["hull_vertices"] = sc["hull_vertices"]; out["hull_offsets"] = sc["hull_offsets"]
        for h, t in enumerate(sc["hull_triangles"]):
            out["hull_triangles%d" % h] = t
    return out


def main():
    for name, (make, steps, iters) in CASES.items():
        sc = make()
        o = oracle_lib.Oracle(sc, canonical=True)
        for _ in range(steps):
            o.step(DT, iters)
        c, b = o.contacts()
        data = pack_scene(sc)
        data.update(steps=np.int32(steps), iterations=np.int32(iters), pairs=o.pairs(), contacts=c, contact_bodies=b, states=o.body_states(), aabbs=o.aabbs())
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, "bodies", len(sc["bodies"]), "pairs", len(o.pairs()), "contacts", len(c), "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
