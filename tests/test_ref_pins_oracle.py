"""Pins the parity oracle (oracle/*.cpp, a restatement) to REFERENCE-OWNED code: the reference's own physics sources
(pkurth/D3D12Renderer src/physics + src/scene + EnTT, the "Physics-Lib" subset) compiled with g++ by oracle/ref_build into
oracle/_ref/libref.so. Both are built with the same floating-point rules (-O2 -ffp-contract=off -fno-fast-math), so wherever the
restatement follows the reference's arithmetic operation for operation the results are BIT-IDENTICAL -- which is what every test
here demands: the reference's physicsStep() on a real EnTT scene against the oracle on the flat arrays, over all collider type
pairs, all six joint kinds with their motors and limits, multi-step trajectories (sweep-and-prune order, Gauss-Seidel order and
persistent broad-phase state included), and the reference's own content builders (humanoid_ragdoll::create, vehicle::create)
against the host mirror's restatement of them. CPU only."""
import numpy as np
import pytest

from d3d12renderer_b200 import abi, scenes, host
import oracle_lib
import ref_lib

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so not built (needs /root/reference: make -C oracle/ref_build)")

DT = np.float32(1.0 / 120.0)
FIELDS = ("position", "rotation", "linearVelocity", "angularVelocity")


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def assert_states_bit_equal(a, b, what=""):
    for f in FIELDS:
        assert bits_equal(a[f], b[f]), "%s: %s differs by %g" % (what, f, float(np.max(np.abs(a[f].astype(np.float64) - b[f]))))


def run_both(sc, steps, iterations=10, ref=None):
    """Reference physicsStep (scalar path: the parity path) vs the oracle in the reference's own SAP / array order."""
    r = ref or ref_lib.RefScene(sc); o = oracle_lib.Oracle(sc, canonical=False)
    assert np.array_equal(r.storage_indices(), np.arange(r.num_bodies())), "rigid body storage order must equal creation order"
    for s in range(steps):
        r.step(DT, iterations); o.step(DT, iterations)
        assert_states_bit_equal(r.body_states(), o.body_states(), "step %d" % s)
    assert ref_lib.load().ref_assert_count(1) == o.narrow_stats()["asserts"]
    return r, o


def test_known_answer_through_the_reference():
    """SURVEY Appendix C, hand-derived: the compiled reference itself lands on those fp32 values, and so does the oracle."""
    sc = scenes.spheres_on_ground(1, mode="rest"); sc["bodies"]["position"][0] = (0, 0.45, 0)
    r, o = run_both(sc, 1)
    s = r.body_states()[0]
    assert abs(float(s["position"][1]) - 0.4549679) < 1e-7 and abs(float(s["linearVelocity"][1]) - 0.59614736) < 1e-7
    d = ref_lib.RefScene(sc).detect()
    assert len(d["contacts"]) == 1 and d["contacts"][0]["friction_restitution"] == 0xB5041999
    assert abs(float(d["contacts"][0]["penetrationDepth"]) - 0.050000012) < 1e-8 and tuple(d["contacts"][0]["normal"]) == (0.0, -1.0, 0.0)


@pytest.mark.parametrize("mode", ["rest", "drop"])
def test_config1_reference_step_equals_oracle(mode):
    run_both(scenes.spheres_on_ground(32, mode=mode), 4)


def test_detection_stages_equal_incl_sap_order():
    """getWorldSpaceColliders -> broadphase -> narrowphase of the reference against the oracle's stages: AABBs, the sweep-and-prune
    PAIR ORDER (swap-remove active list, first-frame axis), contacts and body pairs, all bit for bit."""
    sc = scenes.sphere_pile_settled(8, 8, 8)
    r = ref_lib.RefScene(sc); o = oracle_lib.Oracle(sc, canonical=False)
    for step in range(3):
        d = r.detect(); o.detect()
        op = o.pairs()
        assert bits_equal(d["aabbs"], o.aabbs())
        assert np.array_equal(d["pairs"], np.stack([op["a"], op["b"]], 1)), "SAP pair order differs at step %d" % step
        oc, ob = o.contacts()
        assert len(oc) == len(d["contacts"]) > 1000 and np.array_equal(ob, d["contact_bodies"])
        for f in ("point", "normal", "penetrationDepth", "friction_restitution"):
            assert bits_equal(oc[f], d["contacts"][f])
        r.step(DT, 10); o.step(DT, 10)
    # canonical order is a permutation of the reference's own pair set
    oc = oracle_lib.Oracle(sc, canonical=True); oc.detect(); cp = oc.pairs()
    d = ref_lib.RefScene(sc).detect()
    assert np.array_equal(np.unique(np.sort(d["pairs"], 1), axis=0), np.stack([cp["a"], cp["b"]], 1))


def test_dense_pile_trajectory_equals_oracle():
    run_both(scenes.sphere_pile_settled(8, 8, 8), 5)


def test_all_21_type_pairs_contacts_equal_the_reference():
    """Every intersection(A, B) routine of collision_narrow.cpp (GJK / EPA, clipping, manifold reduction included), the cylinder pairs
    and AABB-AABB among them: world colliders, pair order and every contact bit-identical to the compiled reference; then 3 steps."""
    sc = scenes.contact_zoo(24)
    r = ref_lib.RefScene(sc); o = oracle_lib.Oracle(sc, canonical=False)
    d = r.detect(); o.detect()
    op = o.pairs()
    assert bits_equal(d["aabbs"], o.aabbs()) and np.array_equal(d["pairs"], np.stack([op["a"], op["b"]], 1))
    oc, ob = o.contacts()
    assert len(oc) == len(d["contacts"]) and np.array_equal(ob, d["contact_bodies"])
    for f in ("point", "normal", "penetrationDepth", "friction_restitution"):
        assert bits_equal(oc[f], d["contacts"][f]), f
    po, ko = o.collisions()
    assert np.array_equal(np.stack([po["a"], po["b"]], 1), d["collision_pairs"]) and np.array_equal(ko, d["contact_counts"])
    col = sc["colliders"]
    ta = np.minimum(col["type"][po["a"]], col["type"][po["b"]]).astype(int); tb = np.maximum(col["type"][po["a"]], col["type"][po["b"]]).astype(int)
    for a in range(6):
        for b in range(a, 6):
            assert ((ta == a) & (tb == b)).sum() >= 12, (a, b)
    run_both(sc, 3)


def test_mixed_primitives_pile_equals_oracle():
    run_both(host.build_scene(host.CONFIG_MIXED_PILE, 6, 5, 6, seed=1578123, spacing=0.72), 4)


def test_event_zoo_triggers_and_global_force_equal_oracle():
    """Bodies of all six collider types falling through trigger volumes of all six types under a global force field: the boolean overlap
    tests (collision_narrow.cpp:1593-1689, every type pair with a rigid body on one side) and the trigger enter / leave diff
    (physics.cpp:970-1034) against the compiled reference, 100 steps, states bit-equal.
    Two reference defects are worked around, not reproduced: handleNonCollisionInteractions mirrors the
    rigid-body and trigger indices (`numRigidBodies - 1 - index`, physics.cpp:961,970, although both are forward storage indices,
    scene.h:144-149), so the reference reports trigger T-1-t / body N-1-b where the overlap was (t, b) -- the comparison maps them back;
    and localized force fields never act locally (physics.cpp:773 asks the PARENT entity for a collider_component, which lives on a
    child entity, scene.h:47-48; the local table stays uninitialised arena memory), so only the global field is part of this pin."""
    sc = scenes.event_zoo()
    keep = sc["colliders"]["objectType"] != abi.OBJ_FORCE_FIELD
    sc["colliders"] = np.ascontiguousarray(sc["colliders"][keep]); sc["transforms"] = np.ascontiguousarray(sc["transforms"][:6])
    nb, nt = len(sc["bodies"]), 6
    sc["force_fields"] = np.zeros((0, 3), dtype=np.float32)
    r = ref_lib.RefScene(sc); o = oracle_lib.Oracle(sc, canonical=False); o.set_force_fields(sc["force_fields"], sc["global_force"])
    enters = leaves = 0
    for s in range(100):
        r.step(DT, 10); o.step(DT, 10)
        assert_states_bit_equal(r.body_states(), o.body_states(), "step %d" % s)
        log = r.trigger_log(); oe, ol = o.trigger_events()
        unmirror = lambda a: sorted((nt - 1 - int(t), nb - 1 - int(b)) for t, b in a)
        key = lambda a: sorted((int(t), int(b)) for t, b in a)
        assert unmirror(log[log[:, 2] == 0][:, :2]) == key(oe) and unmirror(log[log[:, 2] == 1][:, :2]) == key(ol), "trigger events differ at step %d" % s
        enters += len(oe); leaves += len(ol)
    assert enters >= 30 and leaves >= 10


def test_ragdoll_joints_equal_oracle():
    """Hinge + cone-twist chains with limits (constraints.cpp:1079-1307,1782-2070), jointed neighbours colliding, ground contacts."""
    run_both(host.build_scene(host.CONFIG_RAGDOLLS, 3, 1, 3, seed=20260924, mode=0, spacing=2.5), 40)


def test_joint_zoo_equals_oracle():
    """Distance / ball / fixed / slider joints and every motor and limit of hinge, cone-twist and slider, 60 steps, 30 iterations."""
    sc = host.build_scene(host.CONFIG_JOINT_ZOO, 2, 1, 1, seed=11, mode=0, spacing=30.0)
    assert {k: len(v[0]) for k, v in sc["joints"].items()} == {0: 6, 1: 6, 2: 2, 3: 6, 4: 6, 5: 4}
    r, o = run_both(sc, 60, iterations=30)
    assert len(o.contacts()[0]) > 0          # the welded box pair reached the ground


def test_vehicle_scene_equals_oracle():
    run_both(host.build_scene(host.CONFIG_VEHICLES, 2, 1, 1, seed=5, mode=1, spacing=8.0), 60, iterations=30)


def test_reference_ragdoll_builder_equals_host_mirror():
    """humanoid_ragdoll::create of the reference (ragdoll.cpp:10-135) against the host mirror's restatement: identical body records
    (poses AND the mass properties the reference computes from the shapes, physics.cpp:1416-1588, rigid_body.cpp:29-81), identical
    collider / joint counts, and -- stepping the reference's scene against the oracle on the mirror's arrays -- identical joints."""
    sc = host.build_scene(host.CONFIG_RAGDOLLS, 1, 1, 1, seed=1, mode=1, spacing=2.5)
    ref_lib.set_hulls({})
    r = ref_lib.RefScene(); r.add_ragdoll((0, 1.25, 0), 0.0); r.add_ground((0, 0, 0), (20, 4, 20), (0, -4, 0))
    assert r.num_bodies() == 14 and r.num_colliders() == len(sc["colliders"]) == 18 and [r.num_joints(k) for k in range(6)] == [0, 0, 0, 6, 7, 0]
    rb = r.bodies()
    for f in rb.dtype.names:
        assert bits_equal(rb[f], sc["bodies"][f]), f
    run_both(sc, 60, iterations=30, ref=r)


def test_reference_vehicle_builder_equals_host_mirror():
    """vehicle::create of the reference (vehicle.cpp:303-497: 16 parts, 87 colliders, 11 hinges with the engine / steering motors,
    4 ball joints, 1 fixed, 1 slider) against the host mirror's restatement, then 90 steps of the reference against the oracle.
    (The reference leaves the physics transforms unplaced, vehicle.cpp:486-496; the mirror reproduces that.)"""
    sc = host.build_scene(host.CONFIG_REFERENCE_VEHICLE, 1, 1, 1)
    ref_lib.set_hulls({})
    r = ref_lib.RefScene(); r.add_vehicle((3, 1, -2), 0.5); r.add_ground((0, 0, 0), (28, 4, 28), (0, -5, 0))
    assert r.num_bodies() == 16 and r.num_colliders() == len(sc["colliders"]) == 87 and [r.num_joints(k) for k in range(6)] == [0, 4, 1, 11, 0, 1]
    rb = r.bodies()
    for f in rb.dtype.names:
        assert bits_equal(rb[f], sc["bodies"][f]), f
    run_both(sc, 90, iterations=30, ref=r)


def test_random_number_generator_and_mass_properties():
    assert np.array_equal(ref_lib.random_floats01(15681923, 4096), scenes.xorshift_floats01(15681923, 4096))      # core/random.h:5-50
    zoo = scenes.contact_zoo(4); ref_lib.set_hulls(zoo)
    orc = oracle_lib.Oracle(zoo)
    for c in zoo["colliders"]:
        i1, c1, m1 = ref_lib.mass_properties(c); i2, c2, m2 = orc.mass_properties(c)
        assert bits_equal(i1, i2) and bits_equal(c1, c2) and np.float32(m1) == np.float32(m2), int(c["type"])


def test_reference_avx2_path_runs_and_is_not_the_parity_path():
    """The reference's AVX2 path (simdBroadPhase / simdNarrowPhase / simdConstraintSolver, the timing baseline of bench.py) runs in the
    g++ build and stays close to the scalar path, but is not bit-equal to it (rsqrt, FMA, re-ordered batches: SURVEY fact 7)."""
    sc = scenes.sphere_pile_settled(8, 8, 8)
    a = ref_lib.RefScene(sc); b = ref_lib.RefScene(sc)
    a.step(DT, 10, simd=False); b.step(DT, 10, simd=True)
    sa, sb = a.body_states(), b.body_states()
    assert np.all(np.isfinite(sb["position"]))
    assert float(np.max(np.abs(sa["position"] - sb["position"]))) < 1e-3
    assert not bits_equal(sa["linearVelocity"], sb["linearVelocity"])


def test_heightmap_collision_equals_oracle():
    """heightmapCollision (heightmap_collision.cpp:509-618) over heightmap_collider_component (terrain/heightmap_collider.*):
    spheres, capsules, AABB colliders (unrotated and rotated into OBBs) and OBB colliders fall onto a rolling 2 x 2 chunk terrain. Height
    lookups, the contact list of the first touching step (triangle order of the min/max mip traversal included) and 40 steps of
    physicsStep agree bit for bit."""
    sc = scenes.terrain_drop(10)
    r = ref_lib.RefScene(sc); o = oracle_lib.Oracle(sc, canonical=False)
    rng = np.random.default_rng(3)
    for x, z in rng.uniform(-17, 17, size=(200, 2)).astype(np.float32):
        assert np.float32(r.heightmap_height_at(x, z)).tobytes() == np.float32(o.heightmap_height_at(x, z)).tobytes()
    touched = 0
    for s in range(40):
        if s in (0, 6, 15, 30):
            d = ref_lib.RefScene(dict(sc, bodies=o.bodies())).detect()
            oo = oracle_lib.Oracle(dict(sc, bodies=o.bodies()), canonical=False); oo.detect()
            oc, ob = oo.contacts(); op, on = oo.collisions()
            assert len(d["contacts"]) == len(oc) and bits_equal(d["contacts"].view(np.uint32), oc.view(np.uint32)), "contacts differ at step %d" % s
            assert np.array_equal(d["contact_counts"], on) and np.array_equal(d["collision_pairs"][:, 0], op["a"])
            assert np.all(d["collision_pairs"][:, 1] == 0xFFFF) and np.all(op["b"] == 0xFFFFFFFF)
            assert np.array_equal(d["contact_bodies"]["rbA"], ob["rbA"]) and np.all(ob["rbB"] == len(sc["bodies"]))
            touched = max(touched, len(oc))
        r.step(DT, 10); o.step(DT, 10)
        assert_states_bit_equal(r.body_states(), o.body_states(), "step %d" % s)
    assert touched > 50, "the scene must actually produce terrain contacts (%d)" % touched
    types = sc["colliders"]["type"]
    assert set(types) == {abi.SPHERE, abi.CAPSULE, abi.AABB, abi.OBB}


def test_cloth_equals_oracle():
    """cloth_component (cloth.cpp: wind, velocity / position / drift Gauss-Seidel passes in constraint creation order, damping,
    moving the fixed row with and without the rigid move, a property change that triggers recalculateProperties): 60 steps, positions and
    velocities bit for bit."""
    class OracleCloth(ref_lib._ClothBase):
        def __init__(self, *a, **kw):
            super().__init__(oracle_lib.load(), "oracle_cloth_", *a, **kw)
    args = (3.0, 4.0, 19, 26, 9.0, 0.6, 0.35, 0.9)
    r = ref_lib.RefCloth(*args); o = OracleCloth(*args)
    rng = np.random.default_rng(2)
    q = np.array([0.0, np.sin(0.2), 0.0, np.cos(0.2)], dtype=np.float32)
    for c in (r, o):
        c.set_fixed(q, (0.5, 6.0, -1.0))
    for step in range(60):
        wind = rng.uniform(-3, 3, size=3).astype(np.float32)
        iters = [(0, 1, 0), (2, 1, 0), (1, 2, 1), (3, 0, 2)][step % 4]
        if step == 20:
            for c in (r, o):
                c.set_properties(12.0, 0.3, 0.2, 1.0)
        if step in (30, 45):
            ang = 0.2 + 0.01 * step; q2 = np.array([0.0, np.sin(ang), 0.0, np.cos(ang)], dtype=np.float32)
            for c in (r, o):
                c.set_fixed(q2, (0.5 + 0.02 * step, 6.0, -1.0), move_rigid=(step == 45))
        for c in (r, o):
            c.wind(wind); c.simulate(*iters, DT)
        rp, rv = r.state(); op, ov = o.state()
        assert bits_equal(rp, op) and bits_equal(rv, ov), "cloth differs at step %d: %g" % (step, float(np.max(np.abs(rp - op))))
    assert float(np.max(np.abs(rv))) > 0.1


def _parse_entity_stream(data):
    """serialized_components in order (serialization_binary.cpp:102-133); returns the physics payloads as plain values."""
    import struct
    off = 0; out = {}
    def take(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, data, off); off += struct.calcsize("<" + fmt); return v
    kinds = ["tag", "transform", "position", "position_rotation", "position_scale", "dynamic_transform", "mesh", "point_light", "spot_light", "rigid_body", "force_field",
             "cloth", "cloth_render", "physics_reference", "terrain", "heightmap_collider", "grass", "proc_placement", "water"]
    used = {0: 4, 1: 7, 2: 7, 3: 6, 4: 10, 5: 8}          # meaningful words of the shape union per collider type
    for kind in kinds:
        (has,) = take("?")
        if not has:
            continue
        if kind == "tag": out[kind] = take("16s")[0].split(b"\0")[0]
        elif kind == "transform": out[kind] = take("10f"); off += 8
        elif kind in ("dynamic_transform", "cloth_render"): out[kind] = True
        elif kind == "rigid_body": out[kind] = take("28f")
        elif kind == "force_field": out[kind] = take("3f")
        elif kind == "physics_reference":
            (nc,) = take("I"); cols = []
            for _ in range(nc):
                words = take("12I"); mat = take("i3f"); (ctype,) = take("B"); off += 15
                cols.append((ctype, (words[:7] + words[8:9]) if ctype == 5 else words[:used[ctype]], mat))       # hull: word 7 is padding in front of the 8-byte index union
            (nj,) = take("I"); joints = []
            for _ in range(nj):
                t, a, b = take("iII"); size = [28, 24, 48, 104, 120, 80][t]
                joints.append((t, a, b, data[off:off + size])); off += size
            out[kind] = (cols, sorted(joints))
        else:
            raise AssertionError("component %s is not part of the physics subset" % kind)
    assert off == len(data)
    return out


def test_binary_entity_stream_layout_and_round_trip():
    """scene/serialization_binary.cpp. The payloads are raw copies of the reference's component structs: their sizes and member
    offsets in the compiled reference are the ones the host mirror's writer / reader assume (and the MSVC sizes SURVEY 8 lists); a stream made
    from the reference's real components of an entity (every collider type) reads back through the host mirror with identical values; entities
    with joints (a ragdoll) survive host write -> host read, and their stream has the same length as the reference's for the same body part."""
    info = np.zeros(16, dtype=np.uint32); ref_lib.load().ref_layout_info(ref_lib._p(info))
    assert list(info) == [48, 112, 80, 48, 64, 66, 16, 28, 24, 48, 104, 120, 80, 16, 4, 4]
    sc = scenes.event_zoo()
    r = ref_lib.RefScene(sc)
    hs = host.HostScene()
    for body in range(0, len(sc["bodies"]), 3):
        data = r.serialize_body_entity(body)
        want = _parse_entity_stream(data)
        assert len(want["physics_reference"][0]) == 1 and "rigid_body" in want and "transform" in want
        if want["physics_reference"][0][0][0] == abi.HULL:
            continue                                    # needs the hull geometry registered in the receiving scene: covered by the pile below
        h = hs.deserialize_entity(data)
        assert h is not None
        got = _parse_entity_stream(hs.serialize_entity(h))
        # adding the colliders recomputes mass and inertia from them, in the reference too (scene.h:59-62); this scene's bodies carry made-up ones
        rb_got, rb_want = got.pop("rigid_body"), want.pop("rigid_body")
        assert rb_got[:3] == rb_want[:3] and rb_got[13:] == rb_want[13:]
        assert got == want, "body %d" % body
    pile = host.build_scene(host.CONFIG_MIXED_PILE, 3, 2, 3, seed=1578123, spacing=0.8)       # mass properties derived from the colliders: everything survives
    rp = ref_lib.RefScene(pile); hp = pile["host_scene"]                  # (it owns the hull geometries the colliders point at)
    for body in range(len(pile["bodies"])):
        data = rp.serialize_body_entity(body)
        h = hp.deserialize_entity(data)
        assert h is not None and _parse_entity_stream(hp.serialize_entity(h)) == _parse_entity_stream(data), "pile body %d" % body
    assert {c[0][0] for c in [(_parse_entity_stream(r.serialize_body_entity(b))["physics_reference"][0]) for b in range(len(sc["bodies"]))]} == {0, 1, 2, 3, 4, 5}
    # joints: the host mirror's stream of every ragdoll body part against the stream made from the reference's own ragdoll (same builder, pinned
    # above): same transform / rigid body / collider payloads, the same constraints (type, struct bytes, partner BODY behind the entity handles)
    rag = host.build_scene(host.CONFIG_RAGDOLLS, 1, 1, 1, seed=1, mode=1, spacing=2.5)["host_scene"]       # mode 1: one unrotated ragdoll at (0, 1.25, 0)
    ref_lib.set_hulls({})
    rr = ref_lib.RefScene(); rr.add_ragdoll((0, 1.25, 0), 0.0); rr.add_ground((0, 0, 0), (20, 4, 20), (0, -4, 0))
    ref_body_of = {rr.entity_of_body(b): b for b in range(14)}; host_body_of = {rag.entity_of_body(b): b for b in range(14)}
    total_joints = 0
    for body in range(14):
        want = _parse_entity_stream(rr.serialize_body_entity(body)); got = _parse_entity_stream(rag.serialize_entity(rag.entity_of_body(body)))
        assert got["transform"] == want["transform"] and got["rigid_body"] == want["rigid_body"] and got["tag"] == want["tag"], "part %d" % body
        assert got["physics_reference"][0] == want["physics_reference"][0], "colliders of part %d" % body
        jw = sorted((t, ref_body_of[a], ref_body_of[b], raw) for t, a, b, raw in want["physics_reference"][1])
        jg = sorted((t, host_body_of[a], host_body_of[b], raw) for t, a, b, raw in got["physics_reference"][1])
        assert jg == jw and len(jg) >= 1, "constraints of part %d" % body
        total_joints += len(jg)
    assert total_joints == 2 * 13           # every joint appears in the stream of both of its bodies
