"""Loads the frozen fixtures of tests/golden/ back into scene dicts."""
import os
import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["config1_rest_64", "config1_drop_64", "sphere_pile_5x4x5", "mixed_pile_4x3x4", "ragdolls_2"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    sc = {"bodies": z["bodies"], "colliders": z["colliders"], "transforms": z["transforms"], "joints": {}}
    for k in range(6):
        if "joints%d" % k in z:
            sc["joints"][k] = (z["joints%d" % k], z["joint_pairs%d" % k])
    if "hull_vertices" in z:
        sc["hull_vertices"] = z["hull_vertices"]; sc["hull_offsets"] = z["hull_offsets"]
        sc["hull_triangles"] = [z["hull_triangles%d" % h] for h in range(len(z["hull_offsets"]) - 1)]
    expect = {k: z[k] for k in ("pairs", "contacts", "contact_bodies", "states", "aabbs")}
    return sc, int(z["steps"]), int(z["iterations"]), expect
