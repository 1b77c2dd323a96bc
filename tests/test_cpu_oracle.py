"""CPU-only tests (no GPU): the parity oracle against hand-derived known answers and independent cross-checks, the host
mirror's scene construction against the oracle, and the C ABI surface of the CUDA library.

The reference ships no tests, golden vectors or fixtures, so parity is pinned by
 (1) closed-form known answers derived from the cited formulas,
 (2) brute-force pair enumeration as an independent check of the sweep-and-prune restatement,
 (3) invariants the reference asserts (depth >= 0, accumulated normal impulse >= 0, |tangent impulse| <= mu * normal impulse),
 (4) frozen oracle outputs under tests/golden/ (made by tests/golden/make_golden.py)."""
import ctypes
import json
import os

import numpy as np
import pytest

from d3d12renderer_b200 import abi, scenes, host
import oracle_lib

DT = np.float32(1.0 / 120.0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def world_collider(ctype, shape, restitution=0.1, friction=0.5):
    c = np.zeros(1, dtype=abi.collider_dtype)
    c["shape"][0, :len(shape)] = shape
    c["type"] = ctype; c["restitution"] = restitution; c["friction"] = friction; c["density"] = 1.0
    c["objectType"] = abi.OBJ_STATIC; c["parentTransform"] = abi.NO_TRANSFORM
    return c


# ---------------------------------------------------------------------------------------------------- known answers
def test_appendix_c_single_sphere_known_answer():
    sc = scenes.spheres_on_ground(1, mode="rest")
    sc["bodies"]["position"][0] = (0, 0.45, 0)
    assert abs(float(sc["bodies"]["invMass"][0]) - 1.9098593) < 1e-6 and abs(float(sc["bodies"]["invInertia"][0, 0]) - 19.098593) < 1e-5
    o = oracle_lib.Oracle(sc)
    o.step(DT, 10)
    c, b = o.contacts()
    assert len(c) == 1 and tuple(b[0]) == (0, 1)
    assert c["friction_restitution"][0] == 0xB5041999
    assert tuple(c["normal"][0]) == (0.0, -1.0, 0.0)
    assert abs(float(c["penetrationDepth"][0]) - 0.050000012) < 1e-9
    assert abs(float(c["point"][0][1]) + 0.025000006) < 1e-9
    s = o.body_states()[0]
    assert abs(float(s["linearVelocity"][1]) - 0.59614736) < 1e-7
    assert abs(float(s["position"][1]) - 0.4549679) < 1e-7
    assert tuple(s["rotation"]) == (0.0, 0.0, 0.0, 1.0) and not np.any(s["angularVelocity"])


def test_free_fall_velocity_known_answer():
    """v.y after force integration + damping (rigid_body.cpp:105-118): -9.81/120 / (1 + 0.4/120)."""
    sc = scenes.spheres_on_ground(1, mode="drop")
    sc["bodies"]["position"][0] = (0, 50.0, 0)
    o = oracle_lib.Oracle(sc); o.step(DT, 10)
    assert abs(float(o.body_states()[0]["linearVelocity"][1]) + 0.08147841) < 1e-8


def test_sphere_sphere_contact_known_answer():
    """Two unit spheres 1.5 apart: normal (1,0,0), depth 0.5, point midway (collision_narrow.cpp:376-396)."""
    o = oracle_lib.Oracle()
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0, 0, 0, 1.0]), world_collider(abi.SPHERE, [1.5, 0, 0, 1.0]))
    assert n == 1
    # equal types swap (collision_narrow.cpp:2374): A is the second collider, so the normal points from (1.5,0,0) to the origin
    assert tuple(normal) == (-1.0, 0.0, 0.0) and abs(pd[0][3] - 0.5) < 1e-7 and np.allclose(pd[0][:3], (0.75, 0, 0))


def test_sphere_inside_box_and_coincident_spheres_degenerate_cases():
    o = oracle_lib.Oracle()
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0, 0, 0, 0.5]), world_collider(abi.AABB, [-1, -1, -1, 1, 1, 1]))
    assert n == 1 and tuple(normal) == (0.0, 1.0, 0.0) and abs(pd[0][3] - 0.5) < 1e-7          # centre inside: up normal, depth = r (:464-467)
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [1, 2, 3, 0.5]), world_collider(abi.SPHERE, [1, 2, 3, 0.25]))
    assert n == 1 and tuple(normal) == (0.0, 1.0, 0.0) and abs(pd[0][3] - 0.75) < 1e-7        # coincident centres (:382-386)


def test_box_box_four_contacts_and_signed_depth_quirk():
    """AABB-AABB always returns 4 contacts with depth = p[minAxis] * sign(d) (collision_narrow.cpp:1092-1110)."""
    o = oracle_lib.Oracle()
    a = world_collider(abi.AABB, [-1, -1, -1, 1, 1, 1]); b = world_collider(abi.AABB, [-0.5, 0.8, -0.5, 0.5, 1.8, 0.5])
    n, normal, pd = o.collide_pair(a, b)
    assert n == 4 and abs(abs(normal[1]) - 1.0) < 1e-7 and np.allclose(np.abs(pd[:, 3]), 0.2, atol=1e-6)
    assert len({float(x) for x in pd[:, 3]}) == 1      # all four share one (signed) depth


def test_capsule_capsule_parallel_two_contacts():
    o = oracle_lib.Oracle()
    n, normal, pd = o.collide_pair(world_collider(abi.CAPSULE, [-1, 0, 0, 1, 0, 0, 0.3]), world_collider(abi.CAPSULE, [-0.5, 0.5, 0, 1.5, 0.5, 0, 0.3]))
    assert n == 2 and np.allclose(np.abs(normal), (0, 1, 0), atol=1e-6) and np.allclose(pd[:, 3], 0.1, atol=1e-6)


def test_gjk_epa_pairs_give_plausible_contacts():
    """capsule / obb / hull pairs through GJK + EPA: depth within the EPA tolerance (0.01, collision_epa.h:139) of the analytic value."""
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 1, 1, 1, seed=3, spacing=1.0)
    o = oracle_lib.Oracle(sc)
    box = world_collider(abi.AABB, [-2, -1, -2, 2, 0, 2])
    n, normal, pd = o.collide_pair(world_collider(abi.CAPSULE, [0, 0.2, 0, 0, 1.0, 0, 0.3]), box)     # vertical capsule 0.1 into the top face
    assert n == 1 and abs(normal[1] + 1.0) < 1e-3 and abs(pd[0][3] - 0.1) < 0.011
    hull = np.zeros(1, dtype=abi.collider_dtype); hull["shape"][0, :7] = (0, 0, 0, 1, 0, 0.35, 0); hull["type"] = abi.HULL
    hull["objectType"] = abi.OBJ_STATIC; hull["parentTransform"] = abi.NO_TRANSFORM    # geometry 0: frustum with bottom face at y = -0.4
    n, normal, pd = o.collide_pair(box, hull)
    assert n == 1 and abs(abs(normal[1]) - 1.0) < 1e-3 and abs(pd[0][3] - 0.05) < 0.011
    assert o.narrow_stats()["asserts"] == 0


# ------------------------------------------------------------------------------------- broad phase cross-checks
@pytest.mark.parametrize("dims", [(6, 5, 7), (10, 3, 4)])
def test_sap_equals_brute_force(dims):
    sc = scenes.sphere_pile_settled(*dims)
    for canonical in (True, False):
        o = oracle_lib.Oracle(sc, canonical=canonical)
        for _ in range(3):          # later steps exercise temporal coherence and the variance-axis switch
            o.step(DT, 4)
            got = o.pairs().copy()
            lo = np.minimum(got["a"], got["b"]); hi = np.maximum(got["a"], got["b"])
            got["a"], got["b"] = lo, hi
            assert np.array_equal(np.sort(got, order=["a", "b"]), np.sort(o.brute_pairs(), order=["a", "b"]))


def test_mixed_scene_sap_equals_brute_force_and_invariants():
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 5, 4, 5, seed=1578123, spacing=0.75)
    o = oracle_lib.Oracle(sc)
    o.step(DT, 10)
    assert np.array_equal(o.pairs(), np.sort(o.brute_pairs(), order=["a", "b"]))
    c, _ = o.contacts()
    assert len(c) > 100
    ns = o.narrow_stats()
    assert ns["asserts"] == 0 and ns["epaOutOfMemory"] == 0 and ns["gjkIterCap"] == 0
    # the device polytope capacities (k_narrow_shapes.cuh) must cover what the oracle needed
    assert ns["epaMaxTriangles"] <= 160 and ns["epaMaxEdges"] <= 240 and ns["epaMaxPoints"] <= 24
    imp = np.zeros((len(c), 2), dtype=np.float32)
    o.lib.oracle_get_contact_impulses(o.ctx, imp.ctypes.data_as(ctypes.c_void_p))
    friction = (c["friction_restitution"] >> 16).astype(np.float32) / np.float32(65535.0)
    assert np.all(imp[:, 0] >= 0.0)                                   # constraints.cpp:3433
    assert np.all(np.abs(imp[:, 1]) <= friction * imp[:, 0] + 1e-4)   # constraints.cpp:3413-3415 (bound uses the previous iteration's normal impulse)


def test_solver_order_replay_is_a_permutation_invariant_for_static_contacts():
    """Config 1 at rest: every contact is sphere-vs-static, so ANY visiting order gives identical bits."""
    sc = scenes.spheres_on_ground(8, mode="rest")
    a = oracle_lib.Oracle(sc); b = oracle_lib.Oracle(sc)
    a.step(DT, 10)
    rng = np.random.default_rng(0)
    b.step(DT, 10, contact_order=rng.permutation(64).astype(np.uint32))
    for f in ("position", "linearVelocity"):
        assert np.array_equal(a.body_states()[f], b.body_states()[f])


# ------------------------------------------------------------------------------------------------ host mirror
def test_host_mass_properties_match_oracle_bitwise():
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 3, 2, 3, seed=5, spacing=1.0)
    o = oracle_lib.Oracle(sc); hs = sc["host_scene"]
    extra = [world_collider(abi.CYLINDER, [0, -0.4, 0.1, 0.2, 0.5, 0, 0.3]), world_collider(abi.CAPSULE, [0.1, 0.2, -0.4, 0, -0.5, 0.3, 0.2])]
    cols = [sc["colliders"][i:i + 1] for i in range(len(sc["colliders"]))] + extra
    seen = set()
    for c in cols:
        hi, hc, hm = hs.mass_properties(c); oi, oc, om = o.mass_properties(c)
        assert np.array_equal(hi, oi) and np.array_equal(hc, oc) and hm == om
        seen.add(int(c["type"][0]))
    assert seen == {0, 1, 2, 3, 4, 5}


def test_host_ragdoll_layout_and_index_conventions():
    sc = host.build_scene(host.CONFIG_RAGDOLLS, 2, 1, 1, seed=1, mode=1, spacing=3.0)
    assert len(sc["bodies"]) == 28 and len(sc["colliders"]) == 2 * 17 + 1           # ragdoll.cpp:36-103 + ground
    assert len(sc["joints"][3][0]) == 12 and len(sc["joints"][4][0]) == 14          # 6 hinges + 7 cone-twists each
    col = sc["colliders"]
    # collider index = reverse creation order (physics.cpp:635-641): the ground (created last) is collider 0
    assert col["objectType"][0] == abi.OBJ_STATIC and col["type"][0] == abi.AABB
    # ... and the very first collider created (torso capsule of ragdoll 0, body 0) is the last one
    assert col["objectIndex"][-1] == 0 and col["type"][-1] == abi.CAPSULE
    # body indices are storage order: joint 0 of kind cone-twist connects torso (0) and head (1)
    assert tuple(sc["joints"][4][1][0]) == (0, 1)
    total_mass = float(np.sum(1.0 / sc["bodies"]["invMass"][:14]))
    assert 60.0 < total_mass < 95.0                                                 # a ~75 kg human (ragdoll.cpp:16)
    # compound torso: 4 capsules -> COG above the hip, inertia matrix symmetric positive
    assert sc["bodies"]["localCOGPosition"][0][1] > 0.05


def test_host_scene_matches_python_generator_for_config1():
    a = scenes.spheres_on_ground(8, mode="rest")
    b = host.build_scene(host.CONFIG_SPHERES_ON_GROUND, 8, 1, 8, mode=1, spacing=1.25)
    assert np.array_equal(a["bodies"]["invMass"], b["bodies"]["invMass"]) and np.array_equal(a["bodies"]["invInertia"], b["bodies"]["invInertia"])
    assert np.allclose(np.sort(a["bodies"]["position"], axis=0), np.sort(b["bodies"]["position"], axis=0))


def test_xorshift_generator_matches_reference_definition():
    """core/random.h:5-50: x ^= x << 13; x ^= x >> 7; x ^= x << 17 on 64 bits; float01 = (uint32)x / UINT_MAX."""
    r = scenes.Xorshift64(15681923)
    x = 15681923
    for _ in range(5):
        x ^= (x << 13) & 0xFFFFFFFFFFFFFFFF; x ^= x >> 7; x ^= (x << 17) & 0xFFFFFFFFFFFFFFFF
        assert r.u64() == x
    v = scenes.xorshift_floats01(15681923, 4)
    assert np.all((v >= 0) & (v <= 1)) and len(set(v.tolist())) == 4


# ---------------------------------------------------------------------------------------------------- C ABI surface
def test_cuda_library_loads_and_exports_every_declared_symbol():
    lib = abi.load_library()
    assert lib.b200phys_version() == 1
    header = open(os.path.join(ROOT, "include", "b200phys.h")).read()
    import re
    declared = sorted(set(re.findall(r"^(?:int|uint32_t|uint64_t|void\*?|const char\*)\s+(b200phys_[a-z_]+)\s*\(", header, flags=re.M)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), "libb200phys.so does not export %s" % name
    assert set(declared) == set(abi.EXPORTS)


def test_struct_sizes_match_header():
    assert abi.body_dtype.itemsize == 140 and abi.body_state_dtype.itemsize == 52 and abi.collider_dtype.itemsize == 64
    assert abi.transform_dtype.itemsize == 32 and abi.contact_dtype.itemsize == 32 and abi.pair_dtype.itemsize == 8
    assert [d.itemsize for d in abi.JOINT_DTYPES] == [28, 24, 48, 104, 120, 80]
    assert ctypes.sizeof(abi.Config) == 16 * 4 and ctypes.sizeof(abi.Stats) == 36 * 4   # 16 / 36 four-byte fields in b200phys.h


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(abi.B200PhysError):
        abi.Stepper(max_bodies=4, max_colliders=4)


def test_product_path_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "d3d12renderer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "oracle_lib", "liboracle", "oracle/", "om_step.h", "om_math.h", "om_shapes.h"):
                    assert needle not in text, "%s references the oracle (%s)" % (os.path.join(dirpath, f), needle)


# ------------------------------------------------------------------------------------------------ frozen fixtures
import golden_util


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_oracle_reproduces_golden_fixtures(name):
    sc, steps, iters, expect = golden_util.load(name)
    o = oracle_lib.Oracle(sc, canonical=True)
    for _ in range(steps):
        o.step(DT, iters)
    assert np.array_equal(o.pairs(), expect["pairs"])
    c, b = o.contacts()
    assert np.array_equal(b, expect["contact_bodies"]) and np.array_equal(c["friction_restitution"], expect["contacts"]["friction_restitution"])
    s = o.body_states()
    for f in ("position", "rotation", "linearVelocity", "angularVelocity"):
        # same compiler flags -> normally bit-identical; 1e-6 leaves room for a different libm (atan2/acos) on another host
        assert np.max(np.abs(s[f].astype(np.float64) - expect["states"][f]) / np.maximum(1.0, np.abs(expect["states"][f]))) <= 1e-6


def test_oracle_collision_events_first_step_all_begin_then_none():
    """event_context starts empty (physics.cpp:946-950): the first step reports every collision as a begin event; an unchanged
    contact set reports nothing; the payload is the manifold average and the pre-solve relative velocity."""
    import oracle_lib
    sc = scenes.spheres_on_ground(4, mode="rest")
    orc = oracle_lib.Oracle(sc, canonical=True)
    orc.step(np.float32(1 / 120.0), 10)
    bp, pl, ep = orc.events()
    col, cnt = orc.collisions()
    assert len(bp) == len(col) == 16 and len(ep) == 0
    assert np.allclose(pl[:, 3:6], pl[0, 3:6]) and abs(abs(pl[0, 4]) - 1.0) < 1e-6        # sphere on floor: vertical normal
    g = np.float32(9.81) / 120
    assert np.allclose(np.abs(pl[:, 7]), g / (1 + 0.4 / 120), rtol=1e-4)                  # one step of gravity (damped), before the solver
    orc.step(np.float32(1 / 120.0), 10)
    bp2, _, ep2 = orc.events()
    assert len(bp2) == 0 and len(ep2) == 0


def test_oracle_trigger_and_force_field_scene_is_pinned():
    """Regression pin of the event-zoo scene in the oracle: counts of trigger enter / leave events and overlapping interactions over
    120 steps, and the sphereVsCylinder quirk (bounding_volumes.cpp:723 compares a squared distance with the radius)."""
    import oracle_lib
    sc = scenes.event_zoo()
    orc = oracle_lib.Oracle(sc, canonical=True)
    orc.set_force_fields(sc["force_fields"], sc["global_force"])
    te = tl = inter = 0
    for _ in range(120):
        orc.step(np.float32(1 / 120.0), 10)
        e, l = orc.trigger_events(); te += len(e); tl += len(l); inter += orc.num_interactions()
        assert len(np.unique(e, axis=0)) == len(e)          # de-duplicated per (trigger, body)
    assert (te, tl, inter) == (92, 51, 7046)
    assert orc.narrow_stats()["asserts"] == 0


def test_scene_yaml_round_trip_and_reference_style_file(tmp_path):
    """Host mirror scene files (physics components of the reference's YAML scenes, serialization_yaml.cpp:74-240,395-520): a saved
    scene loads back to the same bodies / colliders; a hand-written file in the reference's layout (with the renderer keys a real
    scene carries) loads its physics components and ignores the rest."""
    from d3d12renderer_b200 import host
    sc = host.build_scene(host.CONFIG_EVENT_SCENE, 3, 1, 2, seed=5, mode=0, spacing=1.5)
    path = tmp_path / "scene.yaml"
    sc["host_scene"].save_yaml(path)
    back = host.HostScene(); n = back.load_yaml(path)
    a, b = sc["host_scene"].arrays(), back.arrays()
    assert n == 6 + 1 + 3 and len(b["bodies"]) == len(a["bodies"]) == 6          # spheres, ground, trigger / wind / global wind
    for f in ("rotation", "position", "invMass", "invInertia", "localCOGPosition", "gravityFactor", "linearDamping", "angularDamping"):
        assert np.allclose(a["bodies"][f], b["bodies"][f], rtol=1e-6, atol=1e-7), f
    # the trigger callback is code, not data: its collider comes back as a static one; everything else is identical
    ca, cb = a["colliders"], b["colliders"]
    assert len(ca) == len(cb) and np.array_equal(ca["type"], cb["type"]) and np.allclose(ca["shape"], cb["shape"], atol=1e-6)
    assert np.allclose(ca["friction"], cb["friction"]) and np.array_equal(ca["objectType"] == abi.OBJ_RIGID_BODY, cb["objectType"] == abi.OBJ_RIGID_BODY)
    assert (cb["objectType"] == abi.OBJ_FORCE_FIELD).sum() == 1

    ref = tmp_path / "reference_style.yaml"
    ref.write_text(
        "Scene: test\n"
        "Camera:\n  Position: [0, 3, 8]\n  Near plane: 0.1\n"
        "Sun:\n  Color: [1, 0.93, 0.76]\n  Cascades: 3\n"
        "Entities:\n"
        "  - Tag: Ball\n"
        "    Transform:\n      Position: [1, 2.5, -3]\n      Rotation: [0, 0, 0, 1]\n      Scale: [1, 1, 1]\n"
        "    Dynamic: true\n"
        "    Mesh:\n      Handle: 1234\n      Flags: 1\n"
        "    Rigid body:\n      Local COG: [0, 0, 0]\n      Inv mass: 1.9098593\n      Inv inertia: [19.098593, 0, 0, 0, 19.098593, 0, 0, 0, 19.098593]\n"
        "      Gravity factor: 1\n      Linear damping: 0.4\n      Angular damping: 0.4\n"
        "    Colliders:\n"
        "      - Type: Sphere\n        Center: [0, 0, 0]\n        Radius: 0.5\n        Restitution: 0.1\n        Friction: 0.5\n        Density: 1\n"
        "  - Tag: Platform\n"
        "    Transform:\n      Position: [0, -4, 0]\n      Rotation: [0, 0, 0, 1]\n      Scale: [1, 1, 1]\n"
        "    Colliders:\n"
        "      - Type: AABB\n        Min corner: [-30, -4, -30]\n        Max corner: [30, 4, 30]\n        Restitution: 0.1\n        Friction: 1\n        Density: 4\n"
        "      - Type: OBB\n        Center: [0, 5, 0]\n        Radius: [1, 0.5, 2]\n        Rotation: [0, 0.38268343, 0, 0.92387953]\n        Restitution: 0.1\n        Friction: 1\n        Density: 4\n"
        "  - Tag: Lamp\n    Point light:\n      Color: [1, 1, 1]\n      Intensity: 20\n")
    hs = host.HostScene(); assert hs.load_yaml(ref) == 3
    arr = hs.arrays()
    assert len(arr["bodies"]) == 1 and np.allclose(arr["bodies"]["position"][0], (1, 2.5, -3))
    assert abs(float(arr["bodies"]["invMass"][0]) - 1.9098593) < 1e-5             # recomputed from the unit-density sphere: the same value
    col = arr["colliders"]
    assert len(col) == 3 and sorted(col["type"].tolist()) == [abi.SPHERE, abi.AABB, abi.OBB]
    obb = col[col["type"] == abi.OBB][0]
    assert np.allclose(obb["shape"][0:4], (0, 0.38268343, 0, 0.92387953)) and np.allclose(obb["shape"][7:10], (1, 0.5, 2))


def test_locomotion_env_sizes_ranges_reset_state_and_reward():
    """The RL environment of learned_locomotion.{h,cpp} over the host mirror (host/locomotion_env.cpp), the parts that need no GPU:
    state / action sizes (learning_state 13 vec3 + the smoothed action, learning_action 7 cone-twist x 3 + 6 hinge), action limits
    from the ragdoll's joint limits (getLimits :357-383), the state right after a reset (positions relative to the torso COG
    projected to the ground, :113-146) and the reward of the reset pose (all error terms zero: fall * 4, :329-355)."""
    from d3d12renderer_b200 import host
    env = host.LocomotionEnv(4, spacing=5.0)
    assert (env.state_size, env.action_size) == (66, 27)
    lo, hi = env.action_ranges()
    d = np.deg2rad
    # neck: twist 90, swing 50; shoulders: twist 90, swing 130; hips: no swing limit given twist 30 -> (-1 => pi); ankles: twist 20, swing 75
    assert np.allclose(hi[0:3], (d(90), d(50), np.pi), atol=1e-6) and np.allclose(lo[0:3], (-d(90), -d(50), -np.pi), atol=1e-6)
    assert np.allclose(hi[9:12], (d(30), np.pi, np.pi), atol=1e-6)                 # left hip: swingLimit -1 means unlimited
    assert np.allclose(lo[21:27], d([-5, -5, -90, -45, -90, -45]), atol=1e-6) and np.allclose(hi[21:27], d([85, 85, 5, 45, 5, 45]), atol=1e-6)
    s = env.reset()
    assert s.shape == (4, 66) and np.allclose(s, s[0], atol=1e-5)                  # every environment starts in the same pose, in its own frame
    st = s[0]
    assert np.allclose(st[0:3], 0) and np.allclose(st[39:66], 0)                   # at rest, zero smoothed action
    torso_pos, head_pos = st[15:18], st[21:24]
    assert abs(torso_pos[0]) < 1e-6 and abs(torso_pos[2]) < 1e-6 and torso_pos[1] > 1.25      # the frame origin is the torso COG on the ground
    assert abs(head_pos[1] - (1.25 + 0.42 * 1.45)) < 1e-5 and head_pos[1] > 1.0               # head at hip height + 0.42 * 1.45
    left_toes, right_toes = st[3:6], st[9:12]
    assert abs(left_toes[0] + right_toes[0]) < 1e-5 and left_toes[0] < 0 < right_toes[0] and left_toes[2] < 0   # mirrored, toes in front (-z)
    for i in range(4):
        assert abs(env.reward(i) - 4.0) < 5e-3            # acos near 1 turns rounding in the rotation term into ~1e-3
    env.close()


# ------------------------------------------------------------------- more hand-derived known answers (the reference ships none)
def _quat(axis, angle):
    a = np.asarray(axis, dtype=np.float64); a /= np.linalg.norm(a)
    return list(a * np.sin(angle / 2)) + [np.cos(angle / 2)]


def test_sphere_capsule_and_sphere_cylinder_known_answers():
    """sphere-capsule = sphere vs the sphere at the closest point of the capsule axis (collision_narrow.cpp:398-406); sphere-cylinder
    inside the slab 0 <= t <= 1 likewise, beyond the caps against the cap disc (:408-449)."""
    o = oracle_lib.Oracle()
    cap = world_collider(abi.CAPSULE, [0, -1, 0, 0, 1, 0, 0.5])
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0.8, 0.25, 0, 0.5]), cap)
    assert n == 1 and np.allclose(normal, (-1, 0, 0), atol=1e-6) and abs(pd[0][3] - 0.2) < 1e-6        # centres 0.8 apart, radii 0.5 + 0.5
    assert np.allclose(pd[0][:3], (0.4, 0.25, 0), atol=1e-6)                                           # midway between the surfaces
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0, 1.9, 0, 0.5]), cap)                  # beyond the end: the hemisphere
    assert n == 1 and np.allclose(normal, (0, -1, 0), atol=1e-6) and abs(pd[0][3] - 0.1) < 1e-6
    cyl = world_collider(abi.CYLINDER, [0, -1, 0, 0, 1, 0, 0.5])
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0.8, 0.25, 0, 0.5]), cyl)               # side: same as the capsule
    assert n == 1 and np.allclose(normal, (-1, 0, 0), atol=1e-6) and abs(pd[0][3] - 0.2) < 1e-6
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0.2, 1.4, 0, 0.5]), cyl)                # above the flat cap (a capsule would be round here)
    assert n == 1 and np.allclose(normal, (0, -1, 0), atol=1e-5) and abs(pd[0][3] - 0.1) < 1e-5
    assert o.collide_pair(world_collider(abi.SPHERE, [0, 1.9, 0, 0.5]), cyl)[0] == 0                   # 0.9 above the cap: the cylinder has no hemisphere


def test_sphere_obb_is_sphere_aabb_in_the_box_frame():
    """collision_narrow.cpp:480-494: the sphere is rotated into the OBB frame, tested against the AABB, the result rotated back."""
    o = oracle_lib.Oracle()
    q = _quat([0, 0, 1], np.pi / 2)          # box x axis -> world y
    obb = world_collider(abi.OBB, q + [0, 0, 0] + [2.0, 0.5, 1.0])       # world extents: 0.5 along x, 2 along y
    n, normal, pd = o.collide_pair(world_collider(abi.SPHERE, [0.9, 0.3, 0.2, 0.5]), obb)
    assert n == 1 and np.allclose(normal, (-1, 0, 0), atol=1e-6) and abs(pd[0][3] - 0.1) < 1e-6        # face at x = 0.5, sphere reaches 0.4
    assert o.collide_pair(world_collider(abi.SPHERE, [1.01, 0, 0, 0.5]), obb)[0] == 0
    assert o.collide_pair(world_collider(abi.SPHERE, [0, 2.4, 0, 0.5]), obb)[0] == 1                   # the long axis is y in the world


def test_obb_obb_face_contact_is_clipped_to_four_points():
    """obb-obb SAT + clipping (collision_narrow.cpp:1179-1527): a small box resting 0.1 deep on a big one gives the four corners of
    its bottom face, each 0.1 deep, normal along y."""
    o = oracle_lib.Oracle()
    big = world_collider(abi.OBB, [0, 0, 0, 1] + [0, 0, 0] + [2, 1, 2])
    small = world_collider(abi.OBB, _quat([0, 1, 0], 0.3) + [0.2, 1.4, -0.1] + [0.5, 0.5, 0.5])
    n, normal, pd = o.collide_pair(big, small)
    assert n == 4 and abs(abs(normal[1]) - 1.0) < 1e-6 and np.allclose(pd[:, 3], 0.1, atol=1e-5)
    corners = np.array([[sx * 0.5, 0, sz * 0.5] for sx in (-1, 1) for sz in (-1, 1)])
    c, s = np.cos(0.3), np.sin(0.3)
    world = np.stack([c * corners[:, 0] + s * corners[:, 2] + 0.2, np.full(4, 0.95), -s * corners[:, 0] + c * corners[:, 2] - 0.1], 1)
    got = pd[:, :3][np.lexsort((pd[:, 2], pd[:, 0]))]; want = world[np.lexsort((world[:, 2], world[:, 0]))]
    assert np.allclose(got[:, [0, 2]], want[:, [0, 2]], atol=1e-4)                                      # x / z of the contact points = the small box's bottom corners


def test_boolean_overlap_tests_known_answers_and_the_cylinder_quirk():
    """overlapCheck's tests (bounding_volumes.h:301-363, bounding_volumes.cpp:704-733,1079-1199): touching counts as overlapping; the
    sphere-vs-cylinder-cap branch compares a SQUARED distance with the radius (bounding_volumes.cpp:723): a sphere of radius 0.25 whose
    centre is 0.45 above the cap rim region reports an overlap although 0.45 > 0.25, because 0.45^2 = 0.2025 <= 0.25."""
    o = oracle_lib.Oracle()
    S = lambda x, y, z, r: world_collider(abi.SPHERE, [x, y, z, r])
    assert o.overlap_pair(S(0, 0, 0, 1), S(2, 0, 0, 1)) and not o.overlap_pair(S(0, 0, 0, 1), S(2.001, 0, 0, 1))
    box = world_collider(abi.AABB, [-1, -1, -1, 1, 1, 1])
    assert o.overlap_pair(S(1.4, 0, 0, 0.5), box) and not o.overlap_pair(S(1.6, 0, 0, 0.5), box)
    assert o.overlap_pair(S(1.3, 1.3, 0, 0.5), box) and not o.overlap_pair(S(1.4, 1.4, 0, 0.5), box)     # corner region: distance, not a box test
    cyl = world_collider(abi.CYLINDER, [0, -1, 0, 0, 1, 0, 0.5])
    assert o.overlap_pair(S(0.2, 1.45, 0, 0.25), cyl)            # the quirk: 0.45 above the cap, radius 0.25, but 0.2025 <= 0.25
    assert not o.overlap_pair(S(0.2, 1.6, 0, 0.25), cyl)         # 0.36 > 0.25
    assert o.overlap_pair(S(0.7, 0, 0, 0.25), cyl) and not o.overlap_pair(S(0.8, 0, 0, 0.25), cyl)       # the side is an ordinary sphere test
    a = world_collider(abi.OBB, [0, 0, 0, 1] + [0, 0, 0] + [1, 1, 1])
    b = world_collider(abi.OBB, _quat([0, 0, 1], np.pi / 4) + [2.3, 0, 0] + [1, 1, 1])                    # rotated 45 deg: reaches 1.414 towards a
    assert o.overlap_pair(a, b) and not o.overlap_pair(a, world_collider(abi.OBB, _quat([0, 0, 1], np.pi / 4) + [2.5, 0, 0] + [1, 1, 1]))
    cap = world_collider(abi.CAPSULE, [0, -1, 0, 0, 1, 0, 0.25])
    assert o.overlap_pair(cap, world_collider(abi.CAPSULE, [0.4, 0, -1, 0.4, 0, 1, 0.25])) and not o.overlap_pair(cap, world_collider(abi.CAPSULE, [0.6, 0, -1, 0.6, 0, 1, 0.25]))
    assert o.overlap_pair(cap, box) and not o.overlap_pair(world_collider(abi.CAPSULE, [1.3, -1, 0, 1.3, 1, 0, 0.25]), box)   # GJK test


def test_ray_interaction_pushes_the_first_body_hit():
    """testPhysicsInteraction (physics.cpp:555-629) in the host mirror: a vertical ray 0.25 off a resting sphere's centre hits it at
    height sqrt(r^2 - 0.25^2) above the centre; that body gets force = direction * strength and torque = (hit - COG) x force, nobody
    else gets anything; static geometry and misses leave every accumulator untouched."""
    sc = host.build_scene(host.CONFIG_SPHERES_ON_GROUND, 3, 1, 3, seed=1, mode=1, spacing=1.25)
    hs = sc["host_scene"]
    b = sc["bodies"]
    target = 4
    x0, y0, z0 = (float(v) for v in b["position"][target])
    hs.test_physics_interaction([x0 + 0.25, 5.0, z0], [0.0, -1.0, 0.0], 1000.0)
    after = hs.arrays()["bodies"]
    assert np.allclose(after["forceAccumulator"][target], (0, -1000, 0))
    assert np.allclose(after["torqueAccumulator"][target], (0, 0, -250.0), atol=1e-3)        # (0.25, 0.433, 0) x (0, -1000, 0)
    others = np.arange(len(b)) != target
    assert not np.any(after["forceAccumulator"][others]) and not np.any(after["torqueAccumulator"][others])
    hs.test_physics_interaction([50.0, 5.0, 50.0], [0.0, -1.0, 0.0], 1000.0)                  # only the static ground is there
    hs.test_physics_interaction([x0, 5.0, z0], [0.0, 1.0, 0.0], 1000.0)                       # pointing away
    again = hs.arrays()["bodies"]
    assert np.array_equal(again["forceAccumulator"], after["forceAccumulator"])
    # of two spheres in a row the nearer one is pushed
    row = [i for i in range(len(b)) if abs(float(b["position"][i][2]) - z0) < 1e-6]
    first = min(row, key=lambda i: float(b["position"][i][0]))
    hs.test_physics_interaction([-20.0, float(b["position"][first][1]), z0], [1.0, 0.0, 0.0], 10.0)
    final = hs.arrays()["bodies"]
    pushed = np.nonzero(np.any(final["forceAccumulator"] != again["forceAccumulator"], axis=1))[0].tolist()
    assert pushed == [first] and np.allclose(final["forceAccumulator"][first] - again["forceAccumulator"][first], (10, 0, 0))
