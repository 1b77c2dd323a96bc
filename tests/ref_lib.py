"""ctypes binding of oracle/_ref/libref.so -- the reference's OWN physics code (pkurth/D3D12Renderer src/physics, src/scene, EnTT)
compiled with g++ by oracle/ref_build (TEST INFRASTRUCTURE ONLY: the checker of the checker, and bench.py's CPU baseline)."""
import ctypes as C
import os
import numpy as np

from d3d12renderer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref.so")
LIB_FAST = os.path.join(ROOT, "oracle", "_ref", "libref_fast.so")
_libs = {}


def available(fast=False):
    return os.path.exists(LIB_FAST if fast else LIB)


def load(fast=False):
    key = bool(fast)
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(LIB_FAST if fast else LIB)
    vp, u32, f32, i32 = C.c_void_p, C.c_uint32, C.c_float, C.c_int
    lib.ref_scene_create.restype = vp
    lib.ref_scene_destroy.argtypes = [vp]
    lib.ref_set_hulls.argtypes = [vp, vp, u32, vp, vp]
    lib.ref_scene_build.argtypes = [vp, vp, u32, vp, u32, vp, u32]
    lib.ref_scene_get_bodies.argtypes = [vp, vp]
    lib.ref_scene_add_joints.argtypes = [vp, i32, vp, vp, u32]
    lib.ref_scene_set_force_fields.argtypes = [vp, vp, u32, vp]
    lib.ref_scene_add_ragdoll.argtypes = [vp, vp, f32]; lib.ref_scene_add_ragdoll.restype = u32
    lib.ref_scene_add_vehicle.argtypes = [vp, vp, f32, f32, f32]; lib.ref_scene_add_vehicle.restype = u32
    lib.ref_scene_add_ground.argtypes = [vp, vp, vp, vp, f32, f32, f32]
    lib.ref_scene_add_heightmap.argtypes = [vp, u32, f32, vp, f32, f32, f32, vp]
    lib.ref_scene_heightmap_height_at.argtypes = [vp, f32, f32]; lib.ref_scene_heightmap_height_at.restype = f32
    for n in ("ref_scene_num_bodies", "ref_scene_num_colliders"):
        getattr(lib, n).argtypes = [vp]; getattr(lib, n).restype = u32
    lib.ref_scene_num_joints.argtypes = [vp, i32]; lib.ref_scene_num_joints.restype = u32
    lib.ref_scene_set_hinge_motor.argtypes = [vp, u32, i32, f32, f32]
    lib.ref_scene_step.argtypes = [vp, f32, u32, i32, u32, i32, i32, i32]
    lib.ref_scene_last_step_seconds.argtypes = [vp]; lib.ref_scene_last_step_seconds.restype = C.c_double
    lib.ref_scene_get_states.argtypes = [vp, vp]
    lib.ref_scene_trigger_log.argtypes = [vp, vp, u32]; lib.ref_scene_trigger_log.restype = u32
    lib.ref_scene_body_storage_index.argtypes = [vp, u32]; lib.ref_scene_body_storage_index.restype = u32
    lib.ref_scene_detect.argtypes = [vp, i32, i32, vp, vp, vp, u32, vp, vp, vp, vp, u32, vp]
    lib.ref_collide_pair.argtypes = [vp, vp, i32, vp, vp]; lib.ref_collide_pair.restype = i32
    lib.ref_random_floats01.argtypes = [C.c_uint64, vp, u32]
    lib.ref_env_sizes.argtypes = [vp, vp]; lib.ref_env_ranges.argtypes = [vp, vp]
    lib.ref_env_reset.argtypes = [C.c_uint64, f32, i32, vp]; lib.ref_env_update.argtypes = [vp, vp, vp]; lib.ref_env_update.restype = i32
    lib.ref_mass_properties.argtypes = [vp, vp, vp, vp]
    lib.ref_layout_info.argtypes = [vp]; lib.ref_scene_serialize_body_entity.argtypes = [vp, u32, vp, C.c_uint64]; lib.ref_scene_serialize_body_entity.restype = C.c_uint64
    lib.ref_scene_entity_of_body.argtypes = [vp, u32]; lib.ref_scene_entity_of_body.restype = u32
    lib.ref_cloth_create.argtypes = [f32, f32, u32, u32, f32, f32, f32, f32]; lib.ref_cloth_create.restype = vp
    lib.ref_cloth_destroy.argtypes = [vp]; lib.ref_cloth_set_fixed.argtypes = [vp, vp, vp, vp, i32]; lib.ref_cloth_set_properties.argtypes = [vp, f32, f32, f32, f32]
    lib.ref_cloth_wind.argtypes = [vp, vp]; lib.ref_cloth_simulate.argtypes = [vp, u32, u32, u32, f32]; lib.ref_cloth_get.argtypes = [vp, vp, vp]
    lib.ref_assert_count.argtypes = [i32]; lib.ref_assert_count.restype = u32
    _libs[key] = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def set_hulls(scene, fast=False):
    """Hull geometries are process-global in the reference (physics.cpp:14)."""
    lib = load(fast)
    if "hull_vertices" not in scene:
        lib.ref_set_hulls(None, None, 0, None, None); return
    v = np.ascontiguousarray(scene["hull_vertices"], dtype=np.float32); o = np.ascontiguousarray(scene["hull_offsets"], dtype=np.uint32)
    tris = [np.ascontiguousarray(t, dtype=np.uint32) for t in scene["hull_triangles"]]
    to = np.zeros(len(tris) + 1, dtype=np.uint32); to[1:] = np.cumsum([len(t) // 3 for t in tris])
    t = np.concatenate(tris) if tris else np.zeros(0, dtype=np.uint32)
    lib.ref_set_hulls(_p(v), _p(o), len(o) - 1, _p(t), _p(to))


class RefScene:
    """One game_scene of the compiled reference."""

    def __init__(self, scene=None, fast=False):
        self.lib = load(fast); self.fast = fast
        self.ptr = C.c_void_p(self.lib.ref_scene_create())
        if scene is not None:
            self.build(scene)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.ref_scene_destroy(self.ptr); self.ptr = None
        except Exception:
            pass

    def build(self, scene):
        set_hulls(scene, self.fast)
        b = np.ascontiguousarray(scene["bodies"]); t = np.ascontiguousarray(scene["transforms"]); c = np.ascontiguousarray(scene["colliders"])
        self.lib.ref_scene_build(self.ptr, _p(b), len(b), _p(t), len(t), _p(c), len(c))
        for kind, (joints, pairs) in sorted(scene.get("joints", {}).items()):
            j = np.ascontiguousarray(joints); p = np.ascontiguousarray(pairs)
            self.lib.ref_scene_add_joints(self.ptr, kind, _p(j), _p(p), len(j))
        if "heightmap" in scene:
            h = scene["heightmap"]; mc = np.ascontiguousarray(h["min_corner"], dtype=np.float32); hh = np.ascontiguousarray(h["heights"], dtype=np.uint16)
            self.lib.ref_scene_add_heightmap(self.ptr, h["chunks_per_dim"], h["chunk_size"], _p(mc), h["amplitude_scale"], h["restitution"], h["friction"], _p(hh))
        if "force_fields" in scene:
            f = np.ascontiguousarray(scene["force_fields"], dtype=np.float32); g = np.ascontiguousarray(scene.get("global_force", np.zeros(3)), dtype=np.float32)
            self.lib.ref_scene_set_force_fields(self.ptr, _p(f), len(f), _p(g))

    def add_ragdoll(self, hip, rotation=0.0):
        h = np.ascontiguousarray(hip, dtype=np.float32); return self.lib.ref_scene_add_ragdoll(self.ptr, _p(h), rotation)

    def add_vehicle(self, pos, rotation=0.0):
        h = np.ascontiguousarray(pos, dtype=np.float32); return self.lib.ref_scene_add_vehicle(self.ptr, _p(h), rotation, 0.0, 0.0)

    def add_ground(self, center, radius, position, restitution=0.1, friction=1.0, density=4.0):
        c = np.ascontiguousarray(center, dtype=np.float32); r = np.ascontiguousarray(radius, dtype=np.float32); p = np.ascontiguousarray(position, dtype=np.float32)
        self.lib.ref_scene_add_ground(self.ptr, _p(c), _p(r), _p(p), restitution, friction, density)

    def set_hinge_motor(self, index, motor_type, value, max_torque):
        self.lib.ref_scene_set_hinge_motor(self.ptr, index, motor_type, value, max_torque)

    def heightmap_height_at(self, x, z): return float(self.lib.ref_scene_heightmap_height_at(self.ptr, x, z))

    def serialize_body_entity(self, body):
        """the binary entity stream of the entity that owns rigid body `body` (restated framing over the reference's own component structs)"""
        buf = np.zeros(1 << 16, dtype=np.uint8); n = self.lib.ref_scene_serialize_body_entity(self.ptr, body, _p(buf), buf.size)
        return buf[:n].tobytes()

    def entity_of_body(self, body): return self.lib.ref_scene_entity_of_body(self.ptr, body)

    def num_bodies(self): return self.lib.ref_scene_num_bodies(self.ptr)
    def num_colliders(self): return self.lib.ref_scene_num_colliders(self.ptr)
    def num_joints(self, kind): return self.lib.ref_scene_num_joints(self.ptr, kind)

    def step(self, dt, iterations, simd=False, fixed_frame_rate=False, frame_rate=120):
        """physicsStep; fixed_frame_rate False = exactly one physicsStepInternal(dt) (physics.cpp:1404-1412)."""
        s = 1 if simd else 0
        self.lib.ref_scene_step(self.ptr, dt, iterations, 1 if fixed_frame_rate else 0, frame_rate, s, s, s)

    def trigger_log(self):
        """(trigger index, rigid body index, 0 enter / 1 leave) of the last step, in the order the reference fired the callbacks."""
        n = self.lib.ref_scene_trigger_log(self.ptr, None, 0)
        out = np.zeros((n, 3), dtype=np.uint32)
        if n:
            self.lib.ref_scene_trigger_log(self.ptr, _p(out), n)
        return out

    def last_step_seconds(self): return self.lib.ref_scene_last_step_seconds(self.ptr)

    def body_states(self):
        out = np.zeros(self.num_bodies(), dtype=abi.body_state_dtype); self.lib.ref_scene_get_states(self.ptr, _p(out)); return out

    def bodies(self):
        out = np.zeros(self.num_bodies(), dtype=abi.body_dtype); self.lib.ref_scene_get_bodies(self.ptr, _p(out)); return out

    def storage_indices(self):
        return np.array([self.lib.ref_scene_body_storage_index(self.ptr, i) for i in range(self.num_bodies())], dtype=np.uint32)

    def detect(self, simd=False):
        """getWorldSpaceColliders -> broadphase -> narrowphase on the current state -> dict (pairs in the reference's SAP order)."""
        nc = self.num_colliders()
        cap_p = nc * 64 + 5000
        aabbs = np.zeros((nc, 6), dtype=np.float32); world = np.zeros((nc, 16), dtype=np.float32)
        counts = np.zeros(5, dtype=np.uint32)
        s = 1 if simd else 0
        while True:
            pairs = np.zeros((cap_p, 2), dtype=np.uint32); cap_c = 4 * cap_p + 64 * nc
            contacts = np.zeros(cap_c, dtype=abi.contact_dtype); cb = np.zeros(cap_c, dtype=abi.body_pair_dtype)
            cpairs = np.zeros((cap_c, 2), dtype=np.uint32); ccounts = np.zeros(cap_c, dtype=np.uint8)
            # detection mutates the SAP context; callers use a fresh scene per call or accept that
            self.lib.ref_scene_detect(self.ptr, s, s, _p(aabbs), _p(world), _p(pairs), cap_p, _p(contacts), _p(cb), _p(cpairs), _p(ccounts), cap_c, _p(counts))
            if counts[1] <= cap_p:
                break
            cap_p = int(counts[1]) + 16
        return dict(aabbs=aabbs, world=world, pairs=pairs[:counts[1]], contacts=contacts[:counts[3]], contact_bodies=cb[:counts[3]],
                    collision_pairs=cpairs[:counts[2]], contact_counts=ccounts[:counts[2]], num_interactions=int(counts[4]))


def collide_pair(col_a, col_b, simd=False, fast=False):
    lib = load(fast)
    a = np.ascontiguousarray(col_a); b = np.ascontiguousarray(col_b)
    normal = np.zeros(3, dtype=np.float32); pd = np.zeros((4, 4), dtype=np.float32)
    n = lib.ref_collide_pair(_p(a), _p(b), 1 if simd else 0, _p(normal), _p(pd))
    return n, normal, pd[:min(n, 4)]


def random_floats01(seed, n):
    out = np.zeros(n, dtype=np.float32); load().ref_random_floats01(seed, _p(out), n); return out


def mass_properties(collider):
    c = np.ascontiguousarray(collider)
    inertia = np.zeros(9, dtype=np.float32); cog = np.zeros(3, dtype=np.float32); mass = np.zeros(1, dtype=np.float32)
    load().ref_mass_properties(_p(c), _p(inertia), _p(cog), _p(mass))
    return inertia, cog, float(mass[0])


class RefEnv:
    """The reference's RL environment (learned_locomotion.cpp:395-488: resetPhysics / updatePhysics, one ragdoll, 60 Hz, 30 solver iterations),
    process-global like the original DLL."""

    def __init__(self, fast=False):
        self.lib = load(fast)
        a, b = C.c_int(0), C.c_int(0)
        self.lib.ref_env_sizes(C.byref(a), C.byref(b)); self.state_size, self.action_size = a.value, b.value

    def ranges(self):
        lo = np.zeros(self.action_size, dtype=np.float32); hi = np.zeros(self.action_size, dtype=np.float32)
        self.lib.ref_env_ranges(_p(lo), _p(hi)); return lo, hi

    def reset(self, seed=1, push_probability=0.0, simd=False):
        """simd=True is what the reference DLL does (AVX2 path); False = its scalar path, the parity path."""
        out = np.zeros(self.state_size, dtype=np.float32)
        self.lib.ref_env_reset(seed, push_probability, 1 if simd else 0, _p(out)); return out

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32); out = np.zeros(self.state_size, dtype=np.float32); rew = np.zeros(1, dtype=np.float32)
        failed = self.lib.ref_env_update(_p(a), _p(out), _p(rew))
        return out, float(rew[0]), bool(failed)


class _ClothBase:
    """Same calls for the compiled reference's cloth_component and the oracle's restatement (prefix selects the library)."""

    def __init__(self, lib, prefix, width, height, gx, gy, total_mass, stiffness=0.5, damping=0.3, gravity_factor=1.0):
        self.lib = lib; self.prefix = prefix; self.n = gx * gy
        self.ptr = C.c_void_p(getattr(lib, prefix + "create")(width, height, gx, gy, total_mass, stiffness, damping, gravity_factor))

    def __del__(self):
        try:
            getattr(self.lib, self.prefix + "destroy")(self.ptr)
        except Exception:
            pass

    def set_fixed(self, rotation, position, scale=(1, 1, 1), move_rigid=False):
        r = np.ascontiguousarray(rotation, dtype=np.float32); p = np.ascontiguousarray(position, dtype=np.float32); s = np.ascontiguousarray(scale, dtype=np.float32)
        getattr(self.lib, self.prefix + "set_fixed")(self.ptr, _p(r), _p(p), _p(s), 1 if move_rigid else 0)

    def set_properties(self, total_mass, stiffness, damping, gravity_factor):
        getattr(self.lib, self.prefix + "set_properties")(self.ptr, total_mass, stiffness, damping, gravity_factor)

    def wind(self, force):
        f = np.ascontiguousarray(force, dtype=np.float32); getattr(self.lib, self.prefix + "wind")(self.ptr, _p(f))

    def simulate(self, vel_iters, pos_iters, drift_iters, dt):
        getattr(self.lib, self.prefix + "simulate")(self.ptr, vel_iters, pos_iters, drift_iters, dt)

    def state(self):
        p = np.zeros((self.n, 3), dtype=np.float32); v = np.zeros((self.n, 3), dtype=np.float32)
        getattr(self.lib, self.prefix + "get")(self.ptr, _p(p), _p(v)); return p, v


class RefCloth(_ClothBase):
    def __init__(self, *a, **kw):
        super().__init__(load(), "ref_cloth_", *a, **kw)
