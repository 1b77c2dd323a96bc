"""ctypes binding of libb200host.so: the host-side C++ mirror of the reference's physics API (d3d12renderer_b200/host/).

Used by tests and bench.py to build the BASELINE.json scenes with the reference's own scene-construction semantics (mass
properties, collider / body index conventions, ragdoll layout, joint builders) and to drive physicsStep()."""
import ctypes as C
import os
import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200host.so")
_lib = None

CONFIG_SPHERES_ON_GROUND, CONFIG_SPHERE_BOX_DROP, CONFIG_MIXED_PILE, CONFIG_RAGDOLLS, CONFIG_EVENT_SCENE, CONFIG_JOINT_ZOO, CONFIG_VEHICLES, CONFIG_REFERENCE_VEHICLE = 1, 2, 3, 4, 5, 6, 7, 8


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libb200host.so is missing: build it with `make -C d3d12renderer_b200/host`")
    abi.load_library()   # dependency (resolved through rpath as well)
    lib = C.CDLL(LIB_PATH)
    vp, u32 = C.c_void_p, C.c_uint32
    lib.b200host_scene_create.restype = vp
    lib.b200host_scene_destroy.argtypes = [vp]
    lib.b200host_build_config.argtypes = [vp, C.c_int, u32, u32, u32, C.c_uint64, C.c_int, C.c_float]
    lib.b200host_build_mixed_slab.argtypes = [vp, u32, u32, u32, u32, u32, C.c_uint64, C.c_float]
    lib.b200host_add_cloth.argtypes = [vp, C.c_float, C.c_float, u32, u32, C.c_float, C.c_float, C.c_float, C.c_float]; lib.b200host_add_cloth.restype = u32
    lib.b200host_cloth_set_fixed.argtypes = [vp, u32, vp, vp, C.c_int]; lib.b200host_cloth_positions.argtypes = [vp, u32, vp]; lib.b200host_cloth_positions.restype = u32
    lib.b200host_serialize_entity.argtypes = [vp, u32, vp, C.c_uint64]; lib.b200host_serialize_entity.restype = C.c_uint64
    lib.b200host_deserialize_entity.argtypes = [vp, vp, C.c_uint64]; lib.b200host_deserialize_entity.restype = u32
    lib.b200host_entity_of_body.argtypes = [vp, u32]; lib.b200host_entity_of_body.restype = u32; lib.b200host_num_entities.argtypes = [vp]; lib.b200host_num_entities.restype = u32
    lib.b200host_set_heightmap.argtypes = [vp, u32, C.c_float, vp, C.c_float, C.c_float, C.c_float, vp]
    for n in ("b200host_num_bodies", "b200host_num_colliders", "b200host_num_hulls", "b200host_num_hull_vertices"):
        getattr(lib, n).argtypes = [vp]; getattr(lib, n).restype = u32
    lib.b200host_num_joints.argtypes = [vp, C.c_int]; lib.b200host_num_joints.restype = u32
    lib.b200host_num_hull_triangles.argtypes = [vp, u32]; lib.b200host_num_hull_triangles.restype = u32
    lib.b200host_get_bodies.argtypes = [vp, vp]
    lib.b200host_get_colliders.argtypes = [vp, vp, vp]; lib.b200host_get_colliders.restype = u32
    lib.b200host_get_joints.argtypes = [vp, C.c_int, vp, vp]
    lib.b200host_get_hulls.argtypes = [vp, vp, vp]
    lib.b200host_get_hull_triangles.argtypes = [vp, u32, vp]
    lib.b200host_mass_properties.argtypes = [vp, vp, vp, vp, vp]
    lib.b200host_physics_step.argtypes = [vp, C.POINTER(C.c_float), C.c_float, u32, C.c_int, u32, C.c_int]
    lib.b200host_physics_step_with_events.argtypes = [vp, C.POINTER(C.c_float), C.c_float, u32, C.c_int, u32, C.c_int]
    lib.b200host_event_log_counts.argtypes = [C.POINTER(u32)]; lib.b200host_event_log_counts.restype = u32
    lib.b200host_event_log_get.argtypes = [vp, vp, vp]
    lib.b200host_trigger_log.argtypes = [vp, u32, C.c_int]; lib.b200host_trigger_log.restype = u32
    lib.b200host_test_physics_interaction.argtypes = [vp, vp, vp, C.c_float]
    lib.b200host_save_yaml.argtypes = [vp, C.c_char_p]; lib.b200host_load_yaml.argtypes = [vp, C.c_char_p]
    lib.b200host_get_render_transforms.argtypes = [vp, vp]
    lib.b200host_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HostScene:
    """One game_scene of the host mirror."""

    def __init__(self):
        self.lib = load_library()
        self.ptr = C.c_void_p(self.lib.b200host_scene_create())

    def close(self):
        if self.ptr:
            self.lib.b200host_scene_destroy(self.ptr); self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_config(self, config, nx, ny=1, nz=1, seed=1, mode=0, spacing=1.25):
        rc = self.lib.b200host_build_config(self.ptr, config, nx, ny, nz, seed, mode, spacing)
        if rc:
            raise ValueError("unknown config %d" % config)
        return self

    def add_cloth(self, width, height, gx, gy, total_mass, stiffness=0.5, damping=0.3, gravity_factor=1.0):
        return self.lib.b200host_add_cloth(self.ptr, width, height, gx, gy, total_mass, stiffness, damping, gravity_factor)

    def cloth_set_fixed(self, cloth, rotation, position, move_rigid=False):
        r = np.ascontiguousarray(rotation, dtype=np.float32); p = np.ascontiguousarray(position, dtype=np.float32)
        self.lib.b200host_cloth_set_fixed(self.ptr, cloth, _p(r), _p(p), 1 if move_rigid else 0)

    def cloth_positions(self, cloth):
        n = self.lib.b200host_cloth_positions(self.ptr, cloth, None)
        out = np.zeros((n, 3), dtype=np.float32); self.lib.b200host_cloth_positions(self.ptr, cloth, _p(out)); return out

    def serialize_entity(self, entity):
        """serializeEntityToMemory -> bytes"""
        buf = np.zeros(1 << 16, dtype=np.uint8); n = self.lib.b200host_serialize_entity(self.ptr, entity, _p(buf), buf.size)
        return buf[:n].tobytes()

    def deserialize_entity(self, data):
        """deserializeEntityFromMemory into a new entity -> its handle (None on failure)"""
        b = np.frombuffer(data, dtype=np.uint8).copy(); h = self.lib.b200host_deserialize_entity(self.ptr, _p(b), b.size)
        return None if h == 0xFFFFFFFF else h

    def entity_of_body(self, body): return self.lib.b200host_entity_of_body(self.ptr, body)

    def set_heightmap(self, h):
        """heightmap_collider_component of the scene (dict as in scenes.terrain_drop)."""
        mc = np.ascontiguousarray(h["min_corner"], dtype=np.float32); hh = np.ascontiguousarray(h["heights"], dtype=np.uint16)
        self.lib.b200host_set_heightmap(self.ptr, int(h["chunks_per_dim"]), float(h["chunk_size"]), _p(mc), float(h["amplitude_scale"]), float(h["restitution"]), float(h["friction"]), _p(hh))

    def arrays(self):
        """Flat boundary arrays (what physicsStepInternal's gather produces)."""
        lib, s = self.lib, self.ptr
        nb, nc = lib.b200host_num_bodies(s), lib.b200host_num_colliders(s)
        bodies = np.zeros(nb, dtype=abi.body_dtype); lib.b200host_get_bodies(s, _p(bodies))
        colliders = np.zeros(nc, dtype=abi.collider_dtype); transforms = np.zeros(max(1, nc), dtype=abi.transform_dtype)
        nt = lib.b200host_get_colliders(s, _p(colliders), _p(transforms))
        out = {"bodies": bodies, "colliders": colliders, "transforms": np.ascontiguousarray(transforms[:max(1, nt)]), "joints": {}}
        for k in range(6):
            n = lib.b200host_num_joints(s, k)
            if n:
                j = np.zeros(n, dtype=abi.JOINT_DTYPES[k]); p = np.zeros(n, dtype=abi.body_pair_dtype)
                lib.b200host_get_joints(s, k, _p(j), _p(p)); out["joints"][k] = (j, p)
        nh = lib.b200host_num_hulls(s)
        if nh:
            nv = lib.b200host_num_hull_vertices(s)
            verts = np.zeros((nv, 3), dtype=np.float32); offs = np.zeros(nh + 1, dtype=np.uint32)
            lib.b200host_get_hulls(s, _p(verts), _p(offs))
            out["hull_vertices"] = verts; out["hull_offsets"] = offs; out["hull_triangles"] = []
            for h in range(nh):
                t = np.zeros(3 * lib.b200host_num_hull_triangles(s, h), dtype=np.uint32)
                lib.b200host_get_hull_triangles(s, h, _p(t)); out["hull_triangles"].append(t)
        return out

    def mass_properties(self, collider):
        c = np.ascontiguousarray(collider)
        inertia = np.zeros(9, dtype=np.float32); cog = np.zeros(3, dtype=np.float32); mass = np.zeros(1, dtype=np.float32)
        self.lib.b200host_mass_properties(self.ptr, _p(c), _p(inertia), _p(cog), _p(mass))
        return inertia, cog, float(mass[0])

    def physics_step(self, timer, dt, iterations=30, fixed_frame_rate=True, frame_rate=120, exact_order=False):
        t = C.c_float(timer)
        rc = self.lib.b200host_physics_step(self.ptr, C.byref(t), dt, iterations, 1 if fixed_frame_rate else 0, frame_rate, 1 if exact_order else 0)
        if rc:
            raise abi.B200PhysError("physicsStep failed: %d" % rc)
        return t.value

    def physics_step_with_events(self, timer, dt, iterations=30, fixed_frame_rate=True, frame_rate=120, exact_order=False):
        """physicsStep with collision callbacks installed -> (timer, begin entity pairs [n,2], begin payload [n,9], end entity pairs [m,2])
        in the order the callbacks fired."""
        t = C.c_float(timer)
        rc = self.lib.b200host_physics_step_with_events(self.ptr, C.byref(t), dt, iterations, 1 if fixed_frame_rate else 0, frame_rate, 1 if exact_order else 0)
        if rc:
            raise abi.B200PhysError("physicsStep failed: %d" % rc)
        ne = C.c_uint32(0); nb = self.lib.b200host_event_log_counts(C.byref(ne))
        be = np.zeros((nb, 2), dtype=np.uint32); pl = np.zeros((nb, 9), dtype=np.float32); ee = np.zeros((ne.value, 2), dtype=np.uint32)
        self.lib.b200host_event_log_get(_p(be), _p(pl), _p(ee))
        return t.value, be, pl, ee

    def trigger_log(self, clear=True):
        """(trigger entity, other entity, type 0 enter / 1 leave) triples logged by the config-5 trigger callback since the last clear."""
        n = self.lib.b200host_trigger_log(None, 0, 0)
        out = np.zeros((n, 3), dtype=np.uint32)
        self.lib.b200host_trigger_log(_p(out), n, 1 if clear else 0)
        return out

    def test_physics_interaction(self, origin, direction, strength=1000.0):
        """testPhysicsInteraction (physics.cpp:555-629): the first rigid body the ray hits gets direction * strength at the hit point."""
        o = np.ascontiguousarray(origin, dtype=np.float32); d = np.ascontiguousarray(direction, dtype=np.float32)
        self.lib.b200host_test_physics_interaction(self.ptr, _p(o), _p(d), float(strength))

    def save_yaml(self, path):
        """Writes the physics components in the reference's scene-file layout (serialization_yaml.cpp)."""
        if self.lib.b200host_save_yaml(self.ptr, str(path).encode()):
            raise IOError("cannot write %s" % path)

    def load_yaml(self, path):
        """Adds the entities of a reference scene file (physics components only); returns how many were created."""
        n = self.lib.b200host_load_yaml(self.ptr, str(path).encode())
        if n < 0:
            raise IOError("%s is not a scene file" % path)
        return n

    def render_transforms(self):
        out = np.zeros((self.lib.b200host_num_bodies(self.ptr), 7), dtype=np.float32)
        self.lib.b200host_get_render_transforms(self.ptr, _p(out)); return out

    def stats(self):
        st = abi.Stats(); self.lib.b200host_get_stats(self.ptr, C.byref(st)); return st


class LocomotionEnv:
    """Batched version of the reference's RL environment (learned_locomotion.cpp:393-488): E ragdolls, one GPU physics step for all.
    reset() / step(actions [E, 27]) -> states [E, 66], rewards [E], failed [E] (a failed environment must be reset by the caller)."""

    def __init__(self, num_envs, spacing=4.0):
        self.lib = load_library()
        lib = self.lib
        lib.b200env_create.restype = C.c_void_p; lib.b200env_create.argtypes = [C.c_uint32, C.c_float]
        lib.b200env_destroy.argtypes = [C.c_void_p]
        lib.b200env_state_size.restype = C.c_uint32; lib.b200env_action_size.restype = C.c_uint32
        lib.b200env_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.b200env_reward.argtypes = [C.c_void_p, C.c_uint32]; lib.b200env_reward.restype = C.c_float
        lib.b200env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.b200env_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.b200env_set_push_probability.argtypes = [C.c_void_p, C.c_float]
        lib.b200env_set_seed.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        lib.b200env_set_exact_order.argtypes = [C.c_void_p, C.c_int]
        lib.b200env_set_device_mode.argtypes = [C.c_void_p, C.c_int]
        self.num_envs = int(num_envs)
        self.state_size = int(lib.b200env_state_size()); self.action_size = int(lib.b200env_action_size())
        self.ptr = C.c_void_p(lib.b200env_create(self.num_envs, spacing))

    def close(self):
        if self.ptr:
            self.lib.b200env_destroy(self.ptr); self.ptr = None

    def reset(self, index=-1):
        n = self.num_envs if index < 0 else 1
        out = np.zeros((n, self.state_size), dtype=np.float32)
        self.lib.b200env_reset(self.ptr, int(index), _p(out)); return out

    def set_push_probability(self, p):
        """Chance per environment and step of a random sideways push (0.02 in the reference, learned_locomotion.cpp:448)."""
        self.lib.b200env_set_push_probability(self.ptr, float(p))

    def set_device_mode(self, on=True):
        """Everything per environment on the GPU (csrc/k_env.cu): actions up, states / rewards / flags down. Choose before the first step."""
        self.lib.b200env_set_device_mode(self.ptr, 1 if on else 0)

    def set_exact_order(self, on=True):
        """Canonical Gauss-Seidel order (the reference's scalar solver order) instead of the fast colour-sorted one."""
        self.lib.b200env_set_exact_order(self.ptr, 1 if on else 0)

    def set_seed(self, index, seed):
        """Seed of environment `index`'s push generator (the reference's xorshift64)."""
        self.lib.b200env_set_seed(self.ptr, int(index), int(seed))

    def reward(self, index):
        return float(self.lib.b200env_reward(self.ptr, int(index)))

    def action_ranges(self):
        lo = np.zeros(self.action_size, dtype=np.float32); hi = np.zeros(self.action_size, dtype=np.float32)
        self.lib.b200env_ranges(self.ptr, _p(lo), _p(hi)); return lo, hi

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, self.action_size)
        states = np.zeros((self.num_envs, self.state_size), dtype=np.float32); rewards = np.zeros(self.num_envs, dtype=np.float32); failed = np.zeros(self.num_envs, dtype=np.int32)
        rc = self.lib.b200env_step(self.ptr, _p(a), _p(states), _p(rewards), _p(failed))
        if rc:
            raise abi.B200PhysError("physicsStep failed inside the environment: %d" % rc)
        return states, rewards, failed.astype(bool)


def build_mixed_slab(nx_total, ny, nz, ix0, ix1, seed=1578123, spacing=0.8):
    """x-layers [ix0, ix1) of the BASELINE config-3 mixed pile (nx_total x ny x nz) + the static boxes of the whole pile; 'slab' = the x range
    this slab owns (planes half-way to the neighbouring layers), for the multi-GPU decomposition."""
    hs = HostScene()
    hs.lib.b200host_build_mixed_slab(hs.ptr, nx_total, ny, nz, ix0, ix1, seed, spacing)
    sc = hs.arrays(); sc["host_scene"] = hs
    hx = np.float32((nx_total - 1) * spacing * 0.5)
    sc["slab"] = (float(np.float32(ix0) * np.float32(spacing) - hx) - 0.5 * spacing, float(np.float32(ix1 - 1) * np.float32(spacing) - hx) + 0.5 * spacing)
    return sc


def build_scene(config, nx, ny=1, nz=1, seed=1, mode=0, spacing=1.25):
    hs = HostScene().build_config(config, nx, ny, nz, seed, mode, spacing)
    sc = hs.arrays(); sc["host_scene"] = hs
    return sc
