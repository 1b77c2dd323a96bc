"""GPU parity tests: the CUDA stepper (through the C ABI, libb200phys.so) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): broad-phase pair indices bit-exact; positions / velocities within 1e-4 relative after one
step. Because the library is built without FMA contraction and follows the reference's operation order, most quantities
are compared far tighter than that (tolerances are written next to each assert).
"""
import numpy as np
import pytest

from d3d12renderer_b200 import abi, scenes
import oracle_lib

pytestmark = pytest.mark.gpu

DT = np.float32(1.0 / 120.0)


def make_stepper(scene, **kw):
    nb, nc = len(scene["bodies"]), len(scene["colliders"])
    nj = max([len(j[0]) for j in scene.get("joints", {}).values()] + [1])
    st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=max(1, len(scene["transforms"])), max_joints=nj, **kw)
    if "hull_vertices" in scene:
        st.set_hulls(scene["hull_vertices"], scene["hull_offsets"])
    st.upload_bodies(np.ascontiguousarray(scene["bodies"]))
    st.upload_transforms(np.ascontiguousarray(scene["transforms"]))
    st.upload_colliders(np.ascontiguousarray(scene["colliders"]))
    for kind, (joints, pairs) in scene.get("joints", {}).items():
        st.upload_joints(kind, joints, pairs)
    if "heightmap" in scene:
        st.upload_heightmap(scene["heightmap"])
    return st


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def assert_states_close(gpu, ref, tol):
    for f in ("position", "rotation", "linearVelocity", "angularVelocity"):
        e = rel_err(gpu[f], ref[f])
        assert e <= tol, "%s differs: %g > %g" % (f, e, tol)


def run_both(scene, iterations=10, flags=abi.STEP_FAST, steps=1):
    """Steps GPU and oracle; the oracle replays the GPU's solve order (like the reference's AVX2 path, any conflict-free
    re-ordering is a legal Gauss-Seidel order; exact mode uses canonical array order on both sides)."""
    st = make_stepper(scene)
    orc = oracle_lib.Oracle(scene, canonical=True)
    out = None
    for _ in range(steps):
        st.step(DT, iterations, flags)
        stats = st.stats()
        gpu_pairs = st.download_pairs()
        gpu_contacts, gpu_cbodies = st.download_contacts()
        order = st.download_solve_order(6)
        joint_orders = {k: st.download_solve_order(k) for k in scene.get("joints", {})}
        gpu_states = st.download_bodies()
        if flags == abi.STEP_EXACT_ORDER:
            orc.step(DT, iterations)
        else:
            orc.step(DT, iterations, contact_order=order, joint_orders=joint_orders)
        out = dict(stats=stats, gpu_pairs=gpu_pairs, gpu_contacts=gpu_contacts, gpu_cbodies=gpu_cbodies, order=order, gpu_states=gpu_states,
                   ref_pairs=orc.pairs(), ref_contacts=orc.contacts(), ref_states=orc.body_states(), ref_aabbs=orc.aabbs(), gpu_aabbs=st.download_aabbs(),
                   oracle=orc, stepper=st)
    return out


def check_detection(r, contact_tol=0.0):
    assert np.array_equal(r["gpu_aabbs"].view(np.uint32), r["ref_aabbs"].view(np.uint32)), "world AABBs must be bit-identical"
    assert len(r["gpu_pairs"]) == len(r["ref_pairs"])
    assert np.array_equal(r["gpu_pairs"], r["ref_pairs"]), "broad-phase pair list must be bit-exact (canonical order)"
    rc, rb = r["ref_contacts"]
    assert len(r["gpu_contacts"]) == len(rc)
    assert np.array_equal(r["gpu_cbodies"], rb)
    assert np.array_equal(r["gpu_contacts"]["friction_restitution"], rc["friction_restitution"])
    for f in ("point", "normal", "penetrationDepth"):
        e = rel_err(r["gpu_contacts"][f], rc[f])
        assert e <= contact_tol, "contact %s differs: %g" % (f, e)


def test_known_answer_single_sphere():
    """one sphere 0.05 deep in the ground; hand-derived fp32 values."""
    sc = scenes.spheres_on_ground(1, mode="rest")
    sc["bodies"]["position"][0] = (0, 0.45, 0)
    r = run_both(sc)
    check_detection(r)
    c = r["gpu_contacts"][0]
    assert c["friction_restitution"] == 0xB5041999
    assert abs(float(c["penetrationDepth"]) - 0.050000012) < 1e-8 and tuple(c["normal"]) == (0.0, -1.0, 0.0)
    s = r["gpu_states"][0]
    assert abs(float(s["linearVelocity"][1]) - 0.59614736) < 1e-7
    assert abs(float(s["position"][1]) - 0.4549679) < 1e-7
    assert tuple(s["rotation"]) == (0.0, 0.0, 0.0, 1.0)
    assert_states_close(r["gpu_states"], r["ref_states"], 0.0)


@pytest.mark.parametrize("mode", ["rest", "drop"])
def test_config1_spheres_on_ground(mode):
    """BASELINE config 1: 1024 spheres over the ground box, 10 PGS iterations."""
    sc = scenes.spheres_on_ground(32, mode=mode)
    r = run_both(sc, iterations=10)
    check_detection(r)
    if mode == "rest":
        assert r["stats"].numContacts == 1024 and r["stats"].numBroadphaseOverlaps == 1024
    # every contact is sphere-vs-static: order independent, expected bit-identical; the gate is 1e-4
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-6)


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_sphere_pile_order_modes(flags):
    """Dense pile with body-body contacts: fast mode is checked against the oracle replaying the GPU's colour order, exact
    mode against the oracle's canonical Gauss-Seidel order."""
    sc = scenes.sphere_pile_settled(8, 8, 8)
    r = run_both(sc, iterations=10, flags=flags)
    check_detection(r)
    assert r["stats"].numContacts > 1000
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-5)


def test_multi_step_trajectory():
    """20 consecutive steps of a small drop scene stay within the gate (detection exact each step)."""
    sc = scenes.spheres_on_ground(8, mode="rest")
    r = run_both(sc, iterations=10, steps=20)
    check_detection(r)
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def test_pairs_match_brute_force():
    sc = scenes.sphere_pile_settled(10, 6, 10)
    r = run_both(sc, iterations=1)
    brute = r["oracle"].brute_pairs()
    assert np.array_equal(np.sort(brute, order=["a", "b"]), r["gpu_pairs"])


# ------------------------------------------------------------------------------------------------ convex shapes, joints
from d3d12renderer_b200 import host


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_mixed_primitives_pile(flags):
    """BASELINE config 3 in miniature: spheres, capsules, OBBs and convex hulls interpenetrating on a tight lattice. Exercises
    the analytic, SAT + clipping and GJK + EPA routines; narrow phase uses only + - * / sqrt, so contacts must agree exactly."""
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 6, 5, 6, seed=1578123, spacing=0.72)
    r = run_both(sc, iterations=10, flags=flags)
    check_detection(r, contact_tol=1e-6)
    ns = r["oracle"].narrow_stats()
    assert r["stats"].epaOverflows == 0 and ns["epaOutOfMemory"] == 0 and r["stats"].gjkIterationCaps == ns["gjkIterCap"]
    assert r["stats"].numContacts > 300
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def test_mixed_primitives_trajectory():
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 5, 4, 5, seed=99, spacing=0.8)
    r = run_both(sc, iterations=10, steps=8)
    check_detection(r, contact_tol=1e-5)
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_ragdolls_joints_and_contacts(flags):
    """BASELINE config 4 in miniature: humanoid ragdolls (hinge + cone-twist chains, capsule / box colliders) falling on the
    ground; 25 steps so that limits, jointed-neighbour contacts and ground contacts are all active."""
    sc = host.build_scene(host.CONFIG_RAGDOLLS, 3, 1, 3, seed=20260924, mode=0, spacing=2.5)
    assert len(sc["joints"][3][0]) == 9 * 6 and len(sc["joints"][4][0]) == 9 * 7
    r = run_both(sc, iterations=10, flags=flags, steps=25)
    check_detection(r, contact_tol=1e-5)
    assert r["stats"].numContacts > 20
    # joints use atan2 / acos / sin / cos whose device and glibc versions differ in the last ulp: the gate, not bit-equality
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def test_host_mirror_physics_step_matches_abi():
    """physicsStep() of the host mirror (fixed 120 Hz sub-stepping + interpolation, reference physics.cpp:1364-1413) drives the
    same CUDA path: after one frame of exactly one sub-step the physics_transform1 poses equal a direct C-ABI step."""
    sc = host.build_scene(host.CONFIG_SPHERES_ON_GROUND, 8, 1, 8, seed=15681923, mode=1, spacing=1.25)
    hs = sc["host_scene"]
    st = make_stepper(sc)
    st.step(DT, 30, abi.STEP_FAST)
    direct = st.download_bodies()
    timer = hs.physics_step(0.0, float(DT), iterations=30)
    via_host = hs.arrays()["bodies"]
    assert abs(timer) < 1e-6
    assert np.array_equal(via_host["position"], direct["position"]) and np.array_equal(via_host["linearVelocity"], direct["linearVelocity"])
    # interpolation factor 0 after exactly one sub-step: the render transform still shows the OLD pose
    assert np.allclose(hs.render_transforms()[:, 4:7], sc["bodies"]["position"])


def test_host_mirror_heightmap_component():
    """heightmap_collider_component on the host mirror's scene: physicsStep() uploads it and the spheres of config 1 (ground box moved
    far below) come to rest on the terrain exactly as through the C ABI."""
    sc = host.build_scene(host.CONFIG_SPHERES_ON_GROUND, 6, 1, 6, seed=15681923, mode=1, spacing=1.25)
    hm = scenes.terrain_drop(2)["heightmap"]
    hm = dict(hm, min_corner=np.array([-16.0, -3.0, -16.0], dtype=np.float32))
    hs = sc["host_scene"]; hs.set_heightmap(hm)
    st = make_stepper(dict(sc, heightmap=hm), max_contacts=64 * len(sc["colliders"]) + 1024)
    timer = 0.0
    for _ in range(20):
        st.step(DT, 30, abi.STEP_FAST)
        timer = hs.physics_step(timer, float(DT), iterations=30)
    direct = st.download_bodies(); via_host = hs.arrays()["bodies"]
    assert st.stats().numContacts > 0
    assert np.array_equal(via_host["position"], direct["position"]) and np.array_equal(via_host["linearVelocity"], direct["linearVelocity"])


def test_host_mirror_cloth_component():
    """cloth_component in the host mirror's scene: physicsStep() creates it on the device, moves its fixed row, applies the global force
    field as wind and simulates it after the rigid bodies (physics.cpp:1354-1358) -- same particles as the direct C-ABI calls."""
    sc = host.build_scene(host.CONFIG_SPHERES_ON_GROUND, 2, 1, 2, seed=15681923, mode=1, spacing=1.25)
    hs = sc["host_scene"]
    cl = hs.add_cloth(2.0, 3.0, 15, 20, 5.0)
    q = np.array([0.0, 0.0, 0.0, 1.0], dtype=np.float32)
    hs.cloth_set_fixed(cl, q, (0.0, 5.0, 0.0))
    st = make_stepper(sc)
    cid = st.cloth_create(2.0, 3.0, 15, 20, 5.0); st.cloth_set_fixed(cid, q, (0.0, 5.0, 0.0))
    timer = 0.0
    for _ in range(15):
        st.step(DT, 30, abi.STEP_FAST); st.cloth_wind((0.0, 0.0, 0.0)); st.cloth_simulate(0, 1, 0, DT)      # the settings' default iteration counts (physics.h:390-392)
        timer = hs.physics_step(timer, float(DT), iterations=30)
    direct, _ = st.cloth_state(cid)
    assert np.array_equal(hs.cloth_positions(cl), direct) and float(np.max(np.abs(direct[:, 1] - 5.0))) > 0.01


import golden_util


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_gpu_reproduces_golden_fixtures(name):
    """Exact-order mode against the frozen oracle outputs of tests/golden/ (canonical pair order, array-order Gauss-Seidel)."""
    sc, steps, iters, expect = golden_util.load(name)
    st = make_stepper(sc)
    for _ in range(steps):
        st.step(DT, iters, abi.STEP_EXACT_ORDER)
    assert np.array_equal(st.download_pairs(), expect["pairs"])
    assert np.array_equal(st.download_aabbs().view(np.uint32), expect["aabbs"].view(np.uint32))
    c, b = st.download_contacts()
    assert np.array_equal(b, expect["contact_bodies"]) and np.array_equal(c["friction_restitution"], expect["contacts"]["friction_restitution"])
    assert_states_close(st.download_bodies(), expect["states"], 1e-4)


def _sorted_events(pairs, payload=None):
    order = np.lexsort((pairs[:, 1], pairs[:, 0])) if len(pairs) else np.zeros(0, dtype=np.int64)
    return (pairs[order], payload[order]) if payload is not None else pairs[order]


def test_collision_begin_end_events_match_the_oracle():
    """handleCollisionCallbacks (physics.cpp:1037-1178): per step the collider pairs that start / stop colliding, with the manifold-
    averaged contact and the pre-solve relative velocity; device diff of consecutive manifold lists against the oracle's merge."""
    sc = scenes.sphere_box_drop(6, 5, 6, spacing=1.04, base_y=0.55)
    sc["bodies"]["linearVelocity"][:, 1] = -3.0          # hit the floor and each other within a few steps, then bounce
    st = make_stepper(sc); st.enable_events(True)
    orc = oracle_lib.Oracle(sc, canonical=True)
    total_b = total_e = 0
    for step in range(40):
        st.step(DT, 10, abi.STEP_EXACT_ORDER); orc.step(DT, 10)
        b, e = st.download_events()
        ob, opl, oe = orc.events()
        gb, gpl = _sorted_events(np.stack([b["colliderA"], b["colliderB"]], 1).astype(np.uint32), np.concatenate([b["point"], b["normal"], b["relativeVelocity"]], 1))
        rb, rpl = _sorted_events(ob, opl)
        assert np.array_equal(gb, rb), "begin events differ at step %d" % step
        assert np.array_equal(_sorted_events(np.stack([e["a"], e["b"]], 1).astype(np.uint32)), _sorted_events(oe)), "end events differ at step %d" % step
        if len(gb):
            assert np.max(np.abs(gpl - rpl) / np.maximum(1.0, np.abs(rpl))) <= 1e-4
        total_b += len(gb); total_e += len(oe)
    assert total_b > 100 and total_e > 10, (total_b, total_e)     # the scene really produces both kinds


def test_host_mirror_collision_callbacks_fire_in_reference_order():
    """physics_settings::collisionBeginCallback / collisionEndCallback through the host mirror: entity pairs sorted like the reference's
    merge loop, payload equal to the C-ABI events of the same step."""
    sc = host.build_scene(host.CONFIG_SPHERES_ON_GROUND, 6, 1, 6, seed=15681923, mode=1, spacing=1.25)
    hs = sc["host_scene"]
    st = make_stepper(sc); st.enable_events(True)
    timer = 0.0
    seen_begin = 0
    for step in range(3):
        st.step(DT, 30, abi.STEP_FAST)
        b, e = st.download_events()
        timer, be, pl, ee = hs.physics_step_with_events(timer, float(DT), iterations=30)
        assert len(be) == len(b) and len(ee) == len(e)
        order = np.lexsort((b["colliderB"], b["colliderA"]))
        assert np.allclose(pl[:, 0:3], b["point"][order]) and np.allclose(pl[:, 6:9], b["relativeVelocity"][order])
        seen_begin += len(be)
    assert seen_begin >= 36          # every resting sphere starts touching the ground in the first step


def test_triggers_and_force_fields_match_the_oracle():
    """handleNonCollisionInteractions (physics.cpp:952-1035) + overlapCheck (collision_narrow.cpp:1593-1689): bodies of all six collider
    types fall through trigger and force-field volumes of all six types. Per step: the same number of overlapping interactions, the
    same trigger enter / leave events; after the run the same body states (the force fields pushed the same bodies equally)."""
    sc = scenes.event_zoo()
    st = make_stepper(sc); st.upload_force_fields(sc["force_fields"], sc["global_force"])
    orc = oracle_lib.Oracle(sc, canonical=True); orc.set_force_fields(sc["force_fields"], sc["global_force"])
    enters = leaves = 0
    for step in range(100):
        st.step(DT, 10, abi.STEP_EXACT_ORDER); orc.step(DT, 10)
        ge, gl = st.download_trigger_events()
        oe, ol = orc.trigger_events()
        assert st.stats().reserved[3] == orc.num_interactions(), "interaction count differs at step %d" % step
        assert np.array_equal(ge, oe) and np.array_equal(gl, ol), "trigger events differ at step %d" % step
        enters += len(ge); leaves += len(gl)
    assert enters >= 60 and leaves >= 20
    assert_states_close(st.download_bodies(), orc.body_states(), 1e-4)
    s = st.stats()
    assert s.assertCount == 0 and s.gjkIterationCaps == 0


def test_host_mirror_triggers_and_force_fields():
    """trigger_component callbacks and force_field_component (localized + global) through the host mirror's physicsStep: same events
    and the same poses as driving the C ABI directly with the gathered arrays."""
    sc = host.build_scene(host.CONFIG_EVENT_SCENE, 5, 1, 4, seed=77, mode=0, spacing=1.5)
    hs = sc["host_scene"]
    col = sc["colliders"]
    assert (col["objectType"] == abi.OBJ_TRIGGER).sum() == 1 and (col["objectType"] == abi.OBJ_FORCE_FIELD).sum() == 1
    st = make_stepper(sc)
    st.upload_force_fields(np.array([[5.0, 0.0, 0.0], [0.0, 0.0, 0.0]], dtype=np.float32), np.array([0.0, 0.0, 1.0], dtype=np.float32))
    hs.trigger_log(clear=True)
    timer = 0.0; direct_events = []
    for step in range(90):
        st.step(DT, 30, abi.STEP_FAST)
        e, l = st.download_trigger_events()
        direct_events += [(0, int(b), 0) for _, b in e] + [(0, int(b), 1) for _, b in l]
        timer = hs.physics_step(timer, float(DT), iterations=30)
    log = hs.trigger_log()
    n = len(sc["bodies"])
    assert len(log) == len(direct_events) and len(log) >= n            # every sphere enters the slab (and most leave it again)
    # the mirror reports entities: trigger entity is the one created after the ground, bodies are entities 0..n-1 in creation order
    assert sorted((int(o), int(t)) for _, o, t in log) == sorted((b, t) for _, b, t in direct_events)
    assert np.allclose(hs.arrays()["bodies"]["position"], st.download_bodies()["position"], atol=1e-6)
    assert np.all(hs.arrays()["bodies"]["position"][:, 2] > sc["bodies"]["position"][:, 2])      # the global field pushed everything along +z


def test_accumulator_upload_overlapping_the_step_matches_the_oracle():
    """b200phys_upload_forces runs on the context's copy stream and is only awaited right before the force integration; the result
    must equal the oracle stepping bodies whose accumulators were set directly (forces and torques, several steps, re-uploaded
    every step like a game loop does)."""
    sc = scenes.sphere_box_drop(5, 4, 5, spacing=1.05)
    n = len(sc["bodies"])
    rng = np.random.default_rng(5)
    st = make_stepper(sc)
    orc = oracle_lib.Oracle(sc, canonical=True)
    states = st.download_bodies()
    for step in range(6):
        ft = rng.uniform(-30, 30, size=(n, 6)).astype(np.float32)
        b = orc.bodies(); b["forceAccumulator"] = ft[:, 0:3]; b["torqueAccumulator"] = ft[:, 3:6]; orc.set_bodies(b)
        st.upload_body_states(states); st.upload_forces(ft)
        st.step(DT, 10, abi.STEP_EXACT_ORDER); orc.step(DT, 10)
        states = st.download_bodies()
        assert_states_close(states, orc.body_states(), 1e-4)


def test_permuted_order_and_inherited_colours_replayed_by_the_oracle():
    """The two set-up shortcuts of the fast path leave the parity contract untouched: B200PHYS_STEP_PERMUTED (pseudo-random order, no
    colouring) and inherited colours (steps 2..8 after a full colouring) both export the order they used, and the oracle replaying
    that order lands on the same state."""
    sc = scenes.sphere_pile_settled(8, 6, 8)
    r = run_both(sc, iterations=10, flags=abi.STEP_PERMUTED, steps=3)
    check_detection(r)
    assert sorted(r["order"].tolist()) == list(range(len(r["order"])))          # a permutation of the manifolds
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)
    r = run_both(sc, iterations=10, flags=abi.STEP_FAST, steps=11)                  # crosses a re-colouring (every 8 steps)
    check_detection(r)
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def test_capacity_overflow_is_reported_not_fatal():
    """Errors are returned, never trapped: a context created too small for the scene's pair / manifold
    count keeps stepping, reports B200PHYS_E_CAPACITY through get_stats with the overflow counters set, and the bodies stay finite;
    a context that is large enough reports nothing."""
    sc = scenes.sphere_pile_settled(8, 6, 8)
    small = make_stepper(sc, max_pairs=512, max_manifolds=256)
    for _ in range(3):
        small.step(DT, 10, abi.STEP_FAST)
    with pytest.raises(abi.B200PhysError):
        small.stats()
    s = small.stats(check=False)
    assert s.overflowPairs == 1 and s.numBroadphaseOverlaps <= 512
    assert np.all(np.isfinite(small.download_bodies()["position"]))
    ok = make_stepper(sc)
    ok.step(DT, 10, abi.STEP_FAST)
    s = ok.stats()
    assert s.overflowPairs == 0 and s.overflowManifolds == 0 and s.overflowGrid == 0


def test_batched_locomotion_env_steps_on_the_gpu():
    """host/locomotion_env.cpp: 16 ragdoll environments advanced by one physicsStep per call (60 Hz, 30 solver iterations like
    updatePhysics, learned_locomotion.cpp:462-488). Zero actions hold the reset pose with 200 Nm position motors; random actions make
    the figures fall, which shows as failed = head below 1 m; resetting one environment restores its reset state and reward."""
    env = host.LocomotionEnv(16, spacing=5.0)
    env.set_push_probability(0.0)                  # the pose-holding check below wants no random pushes
    s0 = env.reset()
    lo, hi = env.action_ranges()
    states, rewards, failed = env.step(np.zeros((16, 27), dtype=np.float32))
    assert states.shape == (16, 66) and np.all(np.isfinite(states)) and not failed.any()
    assert np.all(rewards > 2.0) and np.all(rewards <= 4.0 + 1e-4)
    assert np.allclose(states[:, 21:24], s0[:, 21:24], atol=0.05)                  # heads have barely moved after 1/60 s
    assert np.abs(states - states[0]).max() < 0.02                                  # same figure in every environment (solver order noise only)
    env.set_push_probability(0.02)                 # the reference's value
    rng = np.random.default_rng(3)
    ever_failed = np.zeros(16, dtype=bool)
    for _ in range(120):
        a = rng.uniform(lo, hi, size=(16, 27)).astype(np.float32)
        states, rewards, failed = env.step(a)
        assert np.all(np.isfinite(states)) and np.all(rewards >= 0.0)
        assert np.array_equal(failed, states[:, 22] < 1.0) and np.all(rewards[failed] == 0.0)
        ever_failed |= failed
    assert ever_failed.sum() >= 8                                                    # flailing figures fall within two seconds
    again = env.reset(3)
    assert np.allclose(again[0], s0[3], atol=1e-5) and abs(env.reward(3) - 4.0) < 5e-3
    env.close()
