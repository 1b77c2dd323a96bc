"""The narrow phase's separation proofs (csrc/k_narrow.cu: spheresApart, sphereBoxApart; csrc/k_bodies.cu: the per-collider bounding sphere)
restated in numpy and checked against the ORACLE on the CPU: over scenes that contain every collider type (spheres, capsules, cylinders,
AABBs incl. the oversized static boxes, OBBs, hulls) no pair the oracle's routines report as colliding may be proved apart, with the
slack the kernels use. The CUDA code itself is compared with the oracle contact by contact in the -m gpu suite; this test pins the
geometry and the slack independently of it, on more samples than the GPU tests step through."""
import numpy as np

from d3d12renderer_b200 import abi, scenes, host
import oracle_lib

DT = np.float32(1.0 / 120.0)


def _qrot(q, v):
    u = q[:3]; t = 2.0 * np.cross(u, v)
    return v + q[3] * t + np.cross(u, t)


def _qmul(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def world_shapes(sc, states):
    """Per collider: bounding sphere (centre, radius) as worldCollidersKernel computes it, and for box-like shapes (world AABB, OBB, hull) the
    box frame (rotation, centre, lo, hi) sphereBoxApart clamps in; None for the others."""
    cols = sc["colliders"]; hv = sc.get("hull_vertices"); ho = sc.get("hull_offsets")
    C = np.zeros((len(cols), 3)); R = np.zeros(len(cols)); boxes = [None] * len(cols)
    ident = np.array([0.0, 0.0, 0.0, 1.0])
    for i, c in enumerate(cols):
        s = c["shape"].astype(np.float64); t = int(c["type"])
        if c["objectType"] == abi.OBJ_RIGID_BODY:
            rot = states["rotation"][c["objectIndex"]].astype(np.float64); pos = states["position"][c["objectIndex"]].astype(np.float64)
        elif c["parentTransform"] != abi.NO_TRANSFORM:
            tr = sc["transforms"][c["parentTransform"]]; rot = tr["rotation"].astype(np.float64); pos = tr["position"].astype(np.float64)
        else:
            rot = ident; pos = np.zeros(3)
        if t == abi.SPHERE:
            C[i] = _qrot(rot, s[0:3]) + pos; R[i] = s[3]
        elif t in (abi.CAPSULE, abi.CYLINDER):
            a = _qrot(rot, s[0:3]) + pos; b = _qrot(rot, s[3:6]) + pos; half = 0.5 * np.linalg.norm(b - a)
            C[i] = 0.5 * (a + b); R[i] = half + s[6] if t == abi.CAPSULE else np.hypot(half, s[6])
        elif t == abi.AABB:
            mn, mx = s[0:3], s[3:6]
            corners = np.array([[x, y, z] for x in (mn[0], mx[0]) for y in (mn[1], mx[1]) for z in (mn[2], mx[2])])
            w = np.array([_qrot(rot, p) + pos for p in corners]); lo, hi = w.min(0), w.max(0)
            C[i] = 0.5 * (lo + hi); R[i] = 0.5 * np.linalg.norm(hi - lo)
            if np.array_equal(rot, ident): boxes[i] = (ident, np.zeros(3), lo, hi)                     # world axes
            else: boxes[i] = (rot, _qrot(rot, 0.5 * (mn + mx)) + pos, -0.5 * (mx - mn), 0.5 * (mx - mn))   # becomes an OBB (physics.cpp:729-733)
        elif t == abi.OBB:
            q = _qmul(rot, s[0:4]); c0 = _qrot(rot, s[4:7]) + pos; C[i] = c0; R[i] = np.linalg.norm(s[7:10]); boxes[i] = (q, c0, -s[7:10], s[7:10])
        else:
            g = int(np.float32(c["shape"][7]).view(np.uint32)); v = hv[ho[g]:ho[g + 1]][:, :3].astype(np.float64)
            q = _qmul(rot, s[0:4]); p = _qrot(rot, s[4:7]) + pos; lo, hi = v.min(0), v.max(0); cl = 0.5 * (lo + hi)
            C[i] = _qrot(q, cl) + p; R[i] = np.linalg.norm(v - cl, axis=1).max() * 1.00001 + 1e-7; boxes[i] = (q, p, lo, hi)
    return C, R, boxes


def spheres_apart(ca, ra, cb, rb):
    reach = (ra + rb) * 1.002 + 1e-4 + 1e-5 * (np.abs(ca).sum() + np.abs(cb).sum())
    return np.dot(ca - cb, ca - cb) > reach * reach


def sphere_box_apart(cs, rs, box):
    q, c, lo, hi = box
    p = _qrot(np.array([-q[0], -q[1], -q[2], q[3]]), cs - c)
    d = p - np.clip(p, lo, hi)
    reach = rs * 1.002 + 1e-4 + 1e-5 * (np.abs(cs).sum() + np.abs(c).sum())
    return np.dot(d, d) > reach * reach


def proved_apart(i, j, C, R, boxes):
    if spheres_apart(C[i], R[i], C[j], R[j]): return True
    if boxes[i] is not None and sphere_box_apart(C[j], R[j], boxes[i]): return True
    if boxes[j] is not None and sphere_box_apart(C[i], R[i], boxes[j]): return True
    return False


def check_scene(sc, steps):
    orc = oracle_lib.Oracle(sc, canonical=True)
    wrong = culled = misses = hits = 0
    for step in range(steps):
        orc.step(DT, 10)
        if step % 3 != steps % 3 and step != steps - 1: continue
        orc.world(); orc.detect()
        pairs = orc.pairs(); col = orc.collisions()[0]
        C, R, boxes = world_shapes(sc, orc.body_states())
        hit = set(map(tuple, np.stack([col["a"], col["b"]], 1).tolist())) | set(map(tuple, np.stack([col["b"], col["a"]], 1).tolist()))
        cols = sc["colliders"]
        for a, b in zip(pairs["a"], pairs["b"]):
            oa, ob = cols["objectType"][a], cols["objectType"][b]
            collision = (oa == abi.OBJ_RIGID_BODY and ob == abi.OBJ_RIGID_BODY and cols["objectIndex"][a] != cols["objectIndex"][b]) or \
                        ((oa == abi.OBJ_STATIC) != (ob == abi.OBJ_STATIC) and abi.OBJ_RIGID_BODY in (oa, ob))
            if not collision: continue
            apart = proved_apart(a, b, C, R, boxes)
            if (int(a), int(b)) in hit:
                hits += 1; wrong += int(apart)
            else:
                misses += 1; culled += int(apart)
    return wrong, culled, misses, hits


def test_no_colliding_pair_is_proved_apart_contact_zoo():
    """all 21 collider type pairs, interpenetrating instances in special and random configurations (cylinders and AABBs included)"""
    wrong, culled, misses, hits = check_scene(scenes.contact_zoo(per_pair=12), steps=4)
    assert hits > 200 and wrong == 0, (wrong, hits)


def test_no_colliding_pair_is_proved_apart_mixed_pile_and_most_misses_are():
    """the benchmark's mixed pile (spheres, capsules, OBBs, hulls inside five oversized static boxes) while it is being thrown around"""
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 10, 12, 10, seed=1578123, spacing=0.8)
    wrong, culled, misses, hits = check_scene(sc, steps=18)
    assert hits > 2000 and wrong == 0, (wrong, hits)
    assert culled > 0.4 * misses, (culled, misses)          # the proofs are worth their cost: more than 40 % of all misses end in the classify kernel


def test_no_colliding_pair_is_proved_apart_ragdolls_and_event_scene():
    """capsule / OBB ragdolls on a ground box; the event scene's triggers, force fields and static colliders with parent transforms"""
    wrong, culled, misses, hits = check_scene(host.build_scene(host.CONFIG_RAGDOLLS, 3, 1, 3, seed=7), steps=40)
    assert hits > 50 and wrong == 0, (wrong, hits)
    wrong, culled, misses, hits = check_scene(scenes.event_zoo(), steps=12)
    assert wrong == 0, (wrong, hits)
