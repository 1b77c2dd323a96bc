"""GPU parity tests, second set (round 2): what round 1 left untested.

  * 16 K - 32 K-body sub-scenes of BASELINE configs 2, 3 and 4 (the grid with many cells, big-collider path, EPA work list, colour
    inheritance at a size where warps, CTAs and scans wrap around), pairs bit-exact, contacts <= 1e-5, states <= 1e-4, every overflow /
    cap counter zero or equal to the oracle's;
  * a contact-level comparison of all 21 collider type pairs, the six cylinder pairs and AABB-AABB included;
  * distance / ball / fixed / slider joints and every motor and limit of hinge, cone-twist and slider; the vehicle acceptance scene;
  * contact / manifold capacity overflow is reported and memory-safe.
All of them call the CUDA path through the C ABI and compare with the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from d3d12renderer_b200 import abi, scenes, host
import oracle_lib
from test_gpu_parity import make_stepper, rel_err, assert_states_close, run_both, check_detection, DT

pytestmark = pytest.mark.gpu


def assert_no_overflow(stats, oracle):
    ns = oracle.narrow_stats()
    assert stats.overflowPairs == 0 and stats.overflowManifolds == 0 and stats.overflowGrid == 0
    assert stats.epaOverflows == 0 and ns["epaOutOfMemory"] == 0
    assert stats.gjkIterationCaps == ns["gjkIterCap"] and stats.assertCount == ns["asserts"]


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_config2_subscene_32k_spheres(flags):
    """BASELINE config 2 at 32 768 spheres (32^3 settled pile in the open box): ~400 k pairs, ~96 k contacts."""
    sc = scenes.sphere_pile_settled(32, 32, 32)
    r = run_both(sc, iterations=10, flags=flags)
    check_detection(r)
    assert r["stats"].numBroadphaseOverlaps > 300000 and r["stats"].numContacts > 80000
    assert_no_overflow(r["stats"], r["oracle"])
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def test_config2_subscene_trajectory_crosses_recolouring():
    """12 consecutive fast-mode steps of a 16 K pile: a full colouring, seven inherited orders, another full colouring, each replayed by the
    oracle; pair list and contacts stay bit-exact every step."""
    sc = scenes.sphere_pile_settled(32, 16, 32)
    st = make_stepper(sc); orc = oracle_lib.Oracle(sc, canonical=True)
    for step in range(12):
        st.step(DT, 10, abi.STEP_FAST)
        order = st.download_solve_order(6)
        orc.step(DT, 10, contact_order=order)
        assert np.array_equal(st.download_pairs(), orc.pairs()), "pairs differ at step %d" % step
        c, b = st.download_contacts(); rc, rb = orc.contacts()
        assert np.array_equal(b, rb) and rel_err(c["penetrationDepth"], rc["penetrationDepth"]) <= 1e-5
        assert_states_close(st.download_bodies(), orc.body_states(), 1e-4)
    assert st.stats().overflowPairs == 0


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_config3_subscene_32k_mixed_primitives(flags):
    """BASELINE config 3 at 32 768 bodies (sphere / capsule / OBB / hull round robin, random orientations, tight lattice): closed-form,
    SAT + clipping and GJK + EPA routines all busy (~130 k manifolds); the EPA work list and its caps at scale."""
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 32, 32, 32, seed=1578123, spacing=0.72)
    r = run_both(sc, iterations=10, flags=flags)
    check_detection(r, contact_tol=1e-5)
    assert r["stats"].numContacts > 100000
    assert_no_overflow(r["stats"], r["oracle"])
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def test_config3_subscene_trajectory():
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 32, 16, 32, seed=77, spacing=0.8)
    r = run_both(sc, iterations=10, steps=4)
    check_detection(r, contact_tol=1e-5)
    assert_no_overflow(r["stats"], r["oracle"])
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


def _orders(st, sc):
    return dict(contact_order=st.download_solve_order(6), joint_orders={k: st.download_solve_order(k) for k in sc.get("joints", {})})


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_config4_subscene_16k_ragdoll_bodies(flags):
    """BASELINE config 4 at 1 156 ragdolls = 16 184 bodies, 19 652 colliders, 8 092 cone-twist + 6 936 hinge joints, packed so
    tightly that neighbouring figures collide. Step 1 from identical inputs: detection bit-exact. Then 28 steps with the <= 1e-4 state
    gate at every step, each from the oracle's state (joints use atan2 / acos / sin / cos whose device and glibc versions differ in the
    last ulp: one step agrees to the gate, not to the bit). The last step's detection is compared bit for bit again: by then feet are on
    the ground, limits are engaged and jointed neighbours collide."""
    sc = host.build_scene(host.CONFIG_RAGDOLLS, 34, 1, 34, seed=20260924, mode=0, spacing=0.7)
    assert len(sc["bodies"]) == 16184 and len(sc["joints"][4][0]) == 1156 * 7 and len(sc["joints"][3][0]) == 1156 * 6
    st = make_stepper(sc); orc = oracle_lib.Oracle(sc, canonical=True)

    def both():
        st.step(DT, 10, flags)
        if flags == abi.STEP_EXACT_ORDER: orc.step(DT, 10)
        else: orc.step(DT, 10, **_orders(st, sc))

    def detection():
        return dict(stats=st.stats(), gpu_pairs=st.download_pairs(), gpu_aabbs=st.download_aabbs(), ref_pairs=orc.pairs(), ref_aabbs=orc.aabbs(),
                    gpu_contacts=st.download_contacts()[0], gpu_cbodies=st.download_contacts()[1], ref_contacts=orc.contacts())
    both()
    check_detection(detection(), contact_tol=1e-6)
    assert_states_close(st.download_bodies(), orc.body_states(), 1e-5)
    st.upload_body_states(orc.body_states())
    for step in range(28):
        both()
        assert_states_close(st.download_bodies(), orc.body_states(), 1e-4)
        st.upload_body_states(orc.body_states())        # the gate is "one step from the same inputs": re-synchronise, so that last-ulp trig
                                                        # differences are not amplified over the (chaotic) 30-step trajectory of 16 k bodies
    both()
    r = detection()
    check_detection(r, contact_tol=1e-5)
    assert r["stats"].numContacts > 2000
    assert_no_overflow(r["stats"], orc)
    assert_states_close(st.download_bodies(), orc.body_states(), 1e-4)


# ----------------------------------------------------------------------------------------- all 21 narrow-phase routines
def test_contact_zoo_every_type_pair_per_contact():
    """Every one of the 21 `intersection(A, B)` routines (collision_narrow.cpp:374-1584) produces manifolds here -- the six cylinder
    pairs and AABB-AABB included -- and every contact (point, normal, depth, packed material, body pair) is compared with the oracle.
    The narrow phase uses + - * / sqrt only and the library is built without FMA contraction: agreement is expected to the last bit
    (tolerance 1e-6 relative as the written bar)."""
    sc = scenes.contact_zoo(per_pair=24)
    r = run_both(sc, iterations=10, flags=abi.STEP_EXACT_ORDER)
    check_detection(r, contact_tol=1e-6)
    pairs, counts = r["stepper"].download_collisions()
    rp, rk = r["oracle"].collisions()
    assert np.array_equal(pairs, rp) and np.array_equal(counts, rk)
    col = sc["colliders"]
    ta = np.minimum(col["type"][pairs["a"]], col["type"][pairs["b"]]).astype(int); tb = np.maximum(col["type"][pairs["a"]], col["type"][pairs["b"]]).astype(int)
    for a in range(6):
        for b in range(a, 6):
            hit = (ta == a) & (tb == b)
            assert hit.sum() >= 12, "type pair (%d, %d) produced only %d manifolds" % (a, b, hit.sum())
    # multi-point manifolds exist for the routines that can produce them
    for a, b in [(1, 1), (1, 2), (2, 2), (1, 3), (2, 3), (3, 3), (3, 4), (4, 4)]:
        assert (counts[(ta == a) & (tb == b)] > 1).any(), (a, b)
    assert_no_overflow(r["stats"], r["oracle"])
    assert_states_close(r["gpu_states"], r["ref_states"], 1e-4)


# ---------------------------------------------------------------------------------------------- joints, motors, vehicle
def _step_compare(sc, steps, flags, tol, iterations=30):
    st = make_stepper(sc); orc = oracle_lib.Oracle(sc, canonical=True)
    for step in range(steps):
        st.step(DT, iterations, flags)
        if flags == abi.STEP_EXACT_ORDER:
            orc.step(DT, iterations)
        else:
            orc.step(DT, iterations, contact_order=st.download_solve_order(6), joint_orders={k: st.download_solve_order(k) for k in sc["joints"]})
        gpu, ref = st.download_bodies(), orc.body_states()
        for f in ("position", "rotation", "linearVelocity", "angularVelocity"):
            e = rel_err(gpu[f], ref[f])
            assert e <= tol, "step %d: %s differs by %g" % (step, f, e)
        assert np.array_equal(st.download_pairs(), orc.pairs()), "pairs differ at step %d" % step
        # the gate is "one step from the same inputs" (joint rows use atan2 / acos / sin / cos, last-ulp differences between the device and
        # glibc; motors and limits amplify them along a trajectory): every step starts from the oracle's state
        st.upload_body_states(ref)
    return st, orc


@pytest.mark.parametrize("flags", [abi.STEP_FAST, abi.STEP_EXACT_ORDER])
def test_joint_zoo_all_kinds_motors_and_limits(flags):
    """Distance, ball, fixed and slider joints (constraints.cpp:189-264,460-528,736-823,2638-2846) plus hinge / cone-twist / slider with
    velocity motors, position motors and limits, hanging from kinematic anchors; 8 copies, 40 steps, compared after EVERY step."""
    sc = host.build_scene(host.CONFIG_JOINT_ZOO, 4, 1, 2, seed=11, mode=0, spacing=30.0)
    counts = {k: len(v[0]) for k, v in sc["joints"].items()}
    assert counts == {0: 24, 1: 24, 2: 8, 3: 24, 4: 24, 5: 16}
    st, orc = _step_compare(sc, 40, flags, 1e-4)
    s = st.stats()
    assert s.assertCount == orc.narrow_stats()["asserts"]
    # the mechanisms really moved: the velocity-motor hinge turned, the slider ran into its limit
    moved = np.abs(st.download_bodies()["position"] - sc["bodies"]["position"]).max(axis=1)
    assert (moved > 0.05).sum() >= 60


def test_vehicle_scene_matches_the_oracle():
    """The reference's vehicle (vehicle.cpp:303-497): 16 parts, 11 hinges (engine velocity motor, steering position motor, suspension
    limits), 1 fixed, 1 slider with limits, 4 ball joints, gear-tooth capsules and wheel cylinders; six of them with the motors running
    drop on the ground and drive for 60 steps."""
    sc = host.build_scene(host.CONFIG_VEHICLES, 3, 1, 2, seed=5, mode=1, spacing=8.0)
    assert len(sc["bodies"]) == 6 * 16 and {k: len(v[0]) for k, v in sc["joints"].items()} == {1: 24, 2: 6, 3: 66, 5: 6}
    st, orc = _step_compare(sc, 60, abi.STEP_FAST, 1e-4)
    c, _ = st.download_contacts(); rc, _ = orc.contacts()
    assert len(c) == len(rc) and len(c) > 20
    assert rel_err(c["point"], rc["point"]) <= 1e-4


# --------------------------------------------------------------------------------------------------- capacity handling
def test_contact_and_manifold_capacity_overflow_is_safe():
    """running out of CONTACT capacity while manifold capacity is left (and vice versa) must cut both lists consistently:
    the step keeps running on a prefix of the canonical manifold list, reports B200PHYS_E_CAPACITY, and every downloaded manifold /
    contact / body pair is a valid prefix of the full result."""
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 8, 6, 8, seed=3, spacing=0.72)
    full = make_stepper(sc); full.step(DT, 10, abi.STEP_EXACT_ORDER)
    fc, fb = full.download_contacts(); fp, fk = full.download_collisions()
    assert len(fc) > 600 and (fk > 1).any()
    for kw in (dict(max_contacts=300), dict(max_manifolds=200), dict(max_contacts=257, max_manifolds=300)):
        small = make_stepper(sc, **kw)
        for _ in range(2):
            small.step(DT, 10, abi.STEP_EXACT_ORDER)
        with pytest.raises(abi.B200PhysError):
            small.stats()
        s = small.stats(check=False)
        assert s.overflowManifolds == 1
        assert np.all(np.isfinite(small.download_bodies()["position"]))
        small2 = make_stepper(sc, **kw); small2.step(DT, 10, abi.STEP_EXACT_ORDER)
        c, b = small2.download_contacts(); p, k = small2.download_collisions()
        assert len(c) <= kw.get("max_contacts", 1 << 30) and len(p) <= kw.get("max_manifolds", 1 << 30)
        assert int(k.sum()) == len(c), "manifold and contact lists must be cut together"
        assert np.array_equal(p, fp[:len(p)]) and np.array_equal(k, fk[:len(k)]) and np.array_equal(b, fb[:len(b)])
        assert np.array_equal(c["point"], fc["point"][:len(c)])


# ------------------------------------------------------------------------------------------------ heightmap (f.4)
@pytest.mark.parametrize("flags", [abi.STEP_EXACT_ORDER, abi.STEP_FAST])
def test_heightmap_collision_matches_oracle(flags):
    """Spheres, capsules, AABB colliders (unrotated and rotated into OBBs) and OBB colliders fall onto a rolling 2 x 2 chunk terrain
    (heightmap_collision.cpp:509-618; the oracle's restatement is pinned bit for bit to the compiled reference by
    tests/test_ref_pins_oracle.py::test_heightmap_collision_equals_oracle). Every step: the contact list (terrain contacts behind the
    body-body ones, one collision {collider, 0xFFFFFFFF} per touching collider, triangle order of the mip traversal) is bit-identical,
    the states agree to 1e-5."""
    sc = scenes.terrain_drop(12)
    st = make_stepper(sc, max_contacts=128 * len(sc["colliders"])); orc = oracle_lib.Oracle(sc, canonical=True)
    most = 0
    for step in range(40):
        st.step(DT, 10, flags)
        s = st.stats()
        assert_no_overflow(s, orc)
        gc, gb = st.download_contacts(); gp, gn = st.download_collisions()
        order = st.download_solve_order(6)
        if flags == abi.STEP_EXACT_ORDER:
            orc.step(DT, 10)
        else:
            orc.step(DT, 10, contact_order=order)
        oc, ob = orc.contacts(); op, on = orc.collisions()
        assert len(gc) == len(oc), "step %d: %d vs %d contacts" % (step, len(gc), len(oc))
        assert np.array_equal(gc.view(np.uint32), oc.view(np.uint32)), "step %d: contacts differ" % step
        assert np.array_equal(gb, ob) and np.array_equal(gn, on) and np.array_equal(gp, op)
        most = max(most, int(np.sum(op["b"] == 0xFFFFFFFF)))
        assert_states_close(st.download_bodies(), orc.body_states(), 1e-5)
        st.upload_body_states(orc.body_states())
    assert most > 100, "terrain collisions per step: %d" % most
    st.close()


# ------------------------------------------------------------------------------------------------ cloth (f.4)
def test_cloth_matches_oracle():
    """cloth_component on the device (csrc/k_cloth.cu: one CTA per cloth, the sequential constraint sweep kept exactly through wavefront
    levels) against the oracle's restatement, which tests/test_ref_pins_oracle.py::test_cloth_equals_oracle pins bit for bit to the compiled
    reference: two cloths (odd and even grid width: both pivot branches of the rigid move), wind, all three solver passes with changing
    iteration counts, a property change (recalculateProperties), moving the fixed row with and without the rigid move. 60 steps, positions
    and velocities bit-identical in every step."""
    import ref_lib

    class OracleCloth(ref_lib._ClothBase):
        def __init__(self, *a, **kw):
            super().__init__(oracle_lib.load(), "oracle_cloth_", *a, **kw)
    sc = scenes.spheres_on_ground(2, mode="rest")
    st = make_stepper(sc)
    specs = [(3.0, 4.0, 19, 26, 9.0, 0.6, 0.35, 0.9), (2.0, 2.5, 24, 17, 4.0, 0.5, 0.3, 1.0)]
    ids = [st.cloth_create(*a) for a in specs]; orc = [OracleCloth(*a) for a in specs]
    rng = np.random.default_rng(2)
    q = np.array([0.0, np.sin(0.2), 0.0, np.cos(0.2)], dtype=np.float32)
    for k in range(2):
        st.cloth_set_fixed(ids[k], q, (0.5 + 3 * k, 6.0, -1.0)); orc[k].set_fixed(q, (0.5 + 3 * k, 6.0, -1.0))
    for step in range(60):
        wind = rng.uniform(-3, 3, size=3).astype(np.float32)
        iters = [(0, 1, 0), (2, 1, 0), (1, 2, 1), (3, 0, 2)][step % 4]
        if step == 20:
            st.cloth_set_properties(ids[0], 12.0, 0.3, 0.2, 1.0); orc[0].set_properties(12.0, 0.3, 0.2, 1.0)
        if step in (30, 45):
            ang = 0.2 + 0.01 * step; q2 = np.array([0.0, np.sin(ang), 0.0, np.cos(ang)], dtype=np.float32)
            for k in range(2):
                st.cloth_set_fixed(ids[k], q2, (0.5 + 3 * k + 0.02 * step, 6.0, -1.0), move_rigid=(step == 45))
                orc[k].set_fixed(q2, (0.5 + 3 * k + 0.02 * step, 6.0, -1.0), move_rigid=(step == 45))
        st.cloth_wind(wind); st.cloth_simulate(*iters, DT)
        for k in range(2):
            orc[k].wind(wind); orc[k].simulate(*iters, DT)
            gp, gv = st.cloth_state(ids[k]); op, ov = orc[k].state()
            assert np.array_equal(gp.view(np.uint32), op.view(np.uint32)) and np.array_equal(gv.view(np.uint32), ov.view(np.uint32)), \
                "cloth %d differs at step %d: %g" % (k, step, float(np.max(np.abs(gp - op))))
    st.close()


# ------------------------------------------------------------------------------------------------ CUDA graph replay
@pytest.mark.parametrize("config,dims,spacing,seed", [(host.CONFIG_MIXED_PILE, (10, 6, 10), 0.8, 1578123), (host.CONFIG_RAGDOLLS, (4, 1, 4), 1.6, 20260924)])
def test_graph_replay_is_bit_identical_to_plain_launches(config, dims, spacing, seed):
    """b200phys_step replays captured CUDA graphs (one per launch-sequence variant: colouring step, inherit steps with either key
    buffer). 40 steps with replay and 40 steps of plain launches from the same state must end in the same bits, and the replay must
    actually have happened."""
    sc = host.build_scene(config, *dims, seed=seed, mode=0, spacing=spacing)
    out = []
    for graphs in (True, False):
        st = make_stepper(sc)
        st.enable_graphs(graphs)
        for _ in range(40):
            st.step(DT, 10, abi.STEP_FAST)
        out.append(st.download_bodies().copy())
        s = st.stats(check=False)
        assert (s.reserved[6] > 20) == graphs, "replays: %d" % s.reserved[6]
        st.close()
    assert out[0].tobytes() == out[1].tobytes()


# ------------------------------------------------------------------------------------------------ RL environment (f.3)
@pytest.mark.parametrize("device_mode", [False, True])
def test_locomotion_env_matches_the_reference_environment(device_mode):
    """The batched environment (host/locomotion_env.cpp over the CUDA step) against the reference's OWN environment, compiled from
    learned_locomotion.cpp (oracle/_ref: resetPhysics / updatePhysics, one ragdoll): same action ranges; with the same random actions
    the 66-float state, the reward and the failure flag agree step by step, also with pushes forced on (probability 1, same generator
    seed: same body part, same ray, same impulse through testPhysicsInteraction)."""
    import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built")
    ref = ref_lib.RefEnv()
    env = host.LocomotionEnv(4, spacing=6.0)
    env.set_device_mode(device_mode)           # the formulas evaluated on the host from downloaded poses, or everything on the GPU (k_env.cu)
    env.set_exact_order(True)                  # joints in array order like the reference's scalar solver (the DLL's default AVX2 path re-orders them)
    assert (ref.state_size, ref.action_size) == (env.state_size, env.action_size) == (66, 27)
    lo, hi = env.action_ranges(); rlo, rhi = ref.ranges()
    assert np.allclose(lo, rlo, atol=1e-6) and np.allclose(hi, rhi, atol=1e-6)
    rng = np.random.default_rng(12)
    # The ragdoll starts 0.16 above the ground and lands on its toes in step 10. Until then the two environments agree to 1e-6 (no
    # actions) / 3e-4 (position motors: device vs glibc trigonometry times the motor gains); once it stands on its feet the stacked
    # contacts amplify that by about 2x per step (measured, tools/env_pin_probe.py: 1e-6 at landing, 5e-4 twelve steps later), and the
    # reference's DLL interface offers no way to re-synchronise. So: (1) no actions, 20 steps = fall, landing and ten steps of standing;
    # (2) random actions for the 8 steps of the fall; (3) the same with a push in every step (same generator seed on both sides).
    # The reward compares looser: it contains acos(w) of rotation differences near identity, where one ulp of w is 3e-4 of angle.
    for scale, push, steps, tol in ((0.0, 0.0, 20, 1e-3), (0.05, 0.0, 8, 1e-3), (0.05, 1.0, 8, 2e-3)):
        env.set_push_probability(push); env.set_seed(0, 77)
        env.reset(); ref.reset(seed=77, push_probability=push, simd=False)
        for step in range(steps):
            a = (scale * rng.uniform(lo, hi)).astype(np.float32)
            states, rewards, failed = env.step(np.tile(a, (4, 1)))
            rs, rr, rf = ref.step(a)
            assert bool(failed[0]) == rf, "failure flag differs at step %d" % step
            err = float(np.max(np.abs(states[0] - rs) / np.maximum(1.0, np.abs(rs))))
            assert err <= tol, "state differs at step %d (action scale %g, push %g): %g" % (step, scale, push, err)
            assert abs(rewards[0] - rr) <= 4e-3, "reward differs at step %d: %g vs %g" % (step, rewards[0], rr)
    env.close()


def test_device_side_environment_matches_the_host_side_one():
    """b200env in device mode (csrc/k_env.cu: action smoothing + motor targets, random pushes, state, reward, failure flag all on the GPU,
    bodies and joints never leave it) against the host-side evaluation of the same formulas (host/locomotion_env.cpp, itself pinned
    against the reference's own environment above): 64 environments, random actions, pushes with probability 0.3 from the same
    generator seeds, resets whenever an environment fails. The physics is the same CUDA step on the same inputs, so the states agree
    to rounding of the few host-vs-device libm calls; the reward goes through acos / exp."""
    E = 64
    envs = []
    for device in (False, True):
        env = host.LocomotionEnv(E, spacing=6.0)
        env.set_device_mode(device)
        env.set_push_probability(0.3)
        for i in range(E):
            env.set_seed(i, 1000 + 17 * i)
        env.reset()
        envs.append(env)
    lo, hi = envs[0].action_ranges()
    rng = np.random.default_rng(5)
    resets = 0
    for step in range(150):
        a = (1.0 * rng.uniform(lo, hi, size=(E, len(lo)))).astype(np.float32)
        sh, rh, fh = envs[0].step(a)
        sd, rd, fd = envs[1].step(a)
        assert np.array_equal(fh, fd), "failure flags differ at step %d" % step
        err = float(np.max(np.abs(sh - sd) / np.maximum(1.0, np.abs(sh))))
        assert err <= 1e-5, "states differ at step %d: %g" % (step, err)
        assert float(np.max(np.abs(rh - rd))) <= 5e-3, "rewards differ at step %d: %g" % (step, float(np.max(np.abs(rh - rd))))
        for i in list(np.nonzero(fh)[0]) + ([3, 7] if step == 25 else []):       # a failed environment is reset; two more once, whatever their state
            envs[0].reset(int(i)); envs[1].reset(int(i)); resets += 1
    assert resets > 2, "no environment ever failed: the reset path is untested"
    for env in envs:
        env.close()


# ------------------------------------------------------------------------------------------------ full-size properties
def _full_size_properties(sc, steps, sample=1500, seed=0):
    """Size-independent checks at a BASELINE size the oracle cannot step in seconds: canonical pair list, soundness of every pair and
    completeness on a random sample of colliders (numpy all-pairs AABB test, the oracle's predicate collision_broad.cpp:316-340), contact
    sanity, counters, and run-to-run determinism."""
    results = []
    for run in range(2):
        st = make_stepper(sc)
        for _ in range(steps):
            st.step(DT, 10, abi.STEP_FAST)
        s = st.stats()
        assert s.overflowPairs == 0 and s.overflowManifolds == 0 and s.overflowGrid == 0 and s.epaOverflows == 0
        pairs = st.download_pairs(); aabbs = st.download_aabbs(); contacts, cb = st.download_contacts(); bodies = st.download_bodies()
        results.append((pairs, contacts, bodies))
        if run == 0:
            a, b = pairs["a"].astype(np.int64), pairs["b"].astype(np.int64)
            assert np.all(a < b), "pairs are stored (low, high)"
            key = a * (1 << 32) + b
            assert np.all(np.diff(key) > 0), "pair list sorted and free of duplicates"
            mn, mx = aabbs[:, :3], aabbs[:, 3:]
            assert np.all(mn[a] <= mx[b]) and np.all(mn[b] <= mx[a]), "every reported pair overlaps (inclusive, like aabbVsAABB)"
            rng = np.random.default_rng(seed)
            pick = rng.choice(len(aabbs), size=sample, replace=False)
            for i in pick:
                hit = np.nonzero(np.all(mn <= mx[i], axis=1) & np.all(mn[i] <= mx, axis=1))[0]
                hit = hit[hit != i]
                lo, hi = np.searchsorted(a, i, "left"), np.searchsorted(a, i, "right")
                mine = set(b[lo:hi].tolist()) | set(a[b == i].tolist())
                assert mine == set(hit.tolist()), "collider %d: %d reported vs %d overlapping" % (i, len(mine), len(hit))
            n = np.linalg.norm(contacts["normal"], axis=1)
            assert len(contacts) == s.numContacts and np.all(np.abs(n - 1.0) < 1e-3) and np.all(contacts["penetrationDepth"] >= 0.0)
            assert np.all(np.isfinite(bodies["position"])) and np.all(np.isfinite(bodies["linearVelocity"]))
        st.close()
    assert np.array_equal(results[0][0], results[1][0]) and results[0][1].tobytes() == results[1][1].tobytes() and results[0][2].tobytes() == results[1][2].tobytes(), "two runs differ"


def test_full_size_config2_properties():
    """BASELINE configs[1] at its full size: 262 144 spheres settling in the open box (5 oversized static boxes: the big-collider path)."""
    _full_size_properties(scenes.sphere_pile_settled(64, 64, 64), steps=6)


def test_full_size_config3_eighth_properties():
    """BASELINE configs[2] lattice at 131 072 mixed primitives (spheres, capsules, OBBs, hulls): grid with two cells per extent, overflow rows,
    every GJK / EPA kernel incl. the redo path."""
    _full_size_properties(host.build_scene(host.CONFIG_MIXED_PILE, 64, 32, 64, seed=1578123, spacing=0.8), steps=6, sample=800)


def _reuse_vs_rebuild(sc, steps, flags=abi.STEP_FAST, kick=None):
    """The same scene in a context that keeps a candidate list between steps and in one that runs the grid pipeline every step: pair
    lists and body states must agree bit for bit in every step (the pair set is a function of indices and AABBs only)."""
    a = make_stepper(sc); b = make_stepper(sc, bp_reuse=False)
    for step in range(steps):
        if kick is not None and step == kick:
            for st in (a, b):      # throw a few bodies across the scene between two steps: their boxes leave the grown ones
                s = st.download_bodies()
                s["position"][::97] += np.array([3.0, 1.5, -2.0], dtype=np.float32); s["linearVelocity"][::97] = np.array([0.0, 4.0, 0.0], dtype=np.float32)
                st.upload_body_states(np.ascontiguousarray(s))
        a.step(DT, 10, flags); b.step(DT, 10, flags)
        pa, pb = a.download_pairs(), b.download_pairs()
        assert np.array_equal(pa, pb), "pair lists differ at step %d (%d vs %d)" % (step, len(pa), len(pb))
        assert a.download_bodies().tobytes() == b.download_bodies().tobytes(), "body states differ at step %d" % step
    ra, rb = a.broadphase_reuse(), b.broadphase_reuse()
    sa, sb = a.stats(), b.stats()
    assert sa.overflowPairs == 0 and sb.overflowPairs == 0 and sa.numBroadphaseOverlaps == sb.numBroadphaseOverlaps
    assert ra["steps"] == steps and rb["steps"] == 0 and rb["margin"] == 0.0
    a.close(); b.close()
    return ra


def test_broadphase_candidate_reuse_resting_scene():
    """1 024 spheres at rest on the ground: after the two opening builds the candidate list survives, the filtered list is the pair list."""
    r = _reuse_vs_rebuild(scenes.spheres_on_ground(32, mode="rest"), steps=30)
    print("reuse (resting spheres):", r)
    assert r["rebuilds"] <= 6 and r["margin"] > 0.0 and r["candidates"] >= r["pairs"]


def test_broadphase_candidate_reuse_settling_pile_with_kick():
    """16 K spheres settling (fast at first: exact builds; slow later: grown boxes), some bodies teleported at step 20: the containment check
    must notice and rebuild; both solve orders."""
    r = _reuse_vs_rebuild(scenes.sphere_pile_settled(32, 16, 32), steps=36, kick=20)
    print("reuse (settling pile, kick at 20):", r)
    assert r["rebuilds"] < 36
    r = _reuse_vs_rebuild(scenes.sphere_pile_settled(16, 16, 16), steps=8, flags=abi.STEP_EXACT_ORDER)


def test_broadphase_candidate_reuse_mixed_pile():
    """8 K mixed primitives (tumbling boxes and hulls change their AABB extents without moving far): reuse vs rebuild contexts bit for bit over
    24 steps; the reuse context against the oracle's sweep-and-prune over the first steps (GPU and oracle states agree to ~1e-6, not bit for bit,
    once GJK / EPA contacts are involved, so a borderline pair may differ later on: the comparison that holds for any number of steps is GPU vs GPU)."""
    sc = host.build_scene(host.CONFIG_MIXED_PILE, 16, 32, 16, seed=1578123, spacing=0.8)
    r = _reuse_vs_rebuild(sc, steps=24)
    print("reuse (mixed pile):", r)
    st = make_stepper(sc); orc = oracle_lib.Oracle(sc, canonical=True)
    for step in range(4):
        st.step(DT, 10, abi.STEP_FAST)
        orc.step(DT, 10, contact_order=st.download_solve_order(6))
        assert np.array_equal(st.download_pairs(), orc.pairs()), "pairs differ at step %d" % step
    assert st.stats().overflowPairs == 0
