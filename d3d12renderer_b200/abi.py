"""ctypes binding of libb200phys.so (include/b200phys.h) + numpy dtypes that mirror the boundary structs.

This is plumbing only: every call forwards to the C ABI. If the CUDA library is missing or cannot be loaded this module
raises -- there is no CPU fallback anywhere in the product path.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200phys.so")

# ---- struct mirrors (include/b200phys.h) --------------------------------------------------------------------------
body_dtype = np.dtype([
    ("rotation", "f4", 4), ("position", "f4", 3), ("localCOGPosition", "f4", 3), ("invMass", "f4"),
    ("invInertia", "f4", 9), ("gravityFactor", "f4"), ("linearDamping", "f4"), ("angularDamping", "f4"),
    ("linearVelocity", "f4", 3), ("angularVelocity", "f4", 3), ("forceAccumulator", "f4", 3), ("torqueAccumulator", "f4", 3)])
body_state_dtype = np.dtype([("rotation", "f4", 4), ("position", "f4", 3), ("linearVelocity", "f4", 3), ("angularVelocity", "f4", 3)])
transform_dtype = np.dtype([("rotation", "f4", 4), ("position", "f4", 3), ("reserved", "f4")])
collider_dtype = np.dtype([
    ("shape", "f4", 10), ("restitution", "f4"), ("friction", "f4"), ("density", "f4"),
    ("type", "u1"), ("objectType", "u1"), ("reserved", "u2"), ("objectIndex", "u4"), ("parentTransform", "u4")])
pair_dtype = np.dtype([("a", "u4"), ("b", "u4")])
body_pair_dtype = np.dtype([("rbA", "u4"), ("rbB", "u4")])
contact_dtype = np.dtype([("point", "f4", 3), ("penetrationDepth", "f4"), ("normal", "f4", 3), ("friction_restitution", "u4")])
begin_event_dtype = np.dtype([("colliderA", "u4"), ("colliderB", "u4"), ("point", "f4", 3), ("normal", "f4", 3), ("relativeVelocity", "f4", 3)])

distance_joint_dtype = np.dtype([("localAnchorA", "f4", 3), ("localAnchorB", "f4", 3), ("globalLength", "f4")])
ball_joint_dtype = np.dtype([("localAnchorA", "f4", 3), ("localAnchorB", "f4", 3)])
fixed_joint_dtype = np.dtype([("initialInvRotationDifference", "f4", 4), ("localAnchorA", "f4", 3), ("localAnchorB", "f4", 3), ("reserved", "f4", 2)])
hinge_joint_dtype = np.dtype([
    ("localAnchorA", "f4", 3), ("localAnchorB", "f4", 3), ("localHingeAxisA", "f4", 3), ("localHingeAxisB", "f4", 3),
    ("minRotationLimit", "f4"), ("maxRotationLimit", "f4"), ("maxMotorTorque", "f4"), ("motorType", "i4"),
    ("motorVelocityOrTargetAngle", "f4"), ("localHingeTangentA", "f4", 3), ("localHingeBitangentA", "f4", 3), ("localHingeTangentB", "f4", 3)])
cone_twist_joint_dtype = np.dtype([
    ("localAnchorA", "f4", 3), ("localAnchorB", "f4", 3), ("localLimitAxisA", "f4", 3), ("localLimitAxisB", "f4", 3),
    ("localLimitTangentA", "f4", 3), ("localLimitBitangentA", "f4", 3), ("localLimitTangentB", "f4", 3),
    ("swingLimit", "f4"), ("twistLimit", "f4"), ("swingMotorType", "i4"), ("swingMotorVelocityOrTargetAngle", "f4"),
    ("maxSwingMotorTorque", "f4"), ("swingMotorAxis", "f4"), ("twistMotorType", "i4"), ("twistMotorVelocityOrTargetAngle", "f4"),
    ("maxTwistMotorTorque", "f4")])
slider_joint_dtype = np.dtype([
    ("initialInvRotationDifference", "f4", 4), ("localAnchorA", "f4", 3), ("localAnchorB", "f4", 3), ("localAxisA", "f4", 3),
    ("negDistanceLimit", "f4"), ("posDistanceLimit", "f4"), ("maxMotorForce", "f4"), ("motorType", "i4"),
    ("motorVelocityOrTargetDistance", "f4"), ("reserved", "f4", 2)])
JOINT_DTYPES = [distance_joint_dtype, ball_joint_dtype, fixed_joint_dtype, hinge_joint_dtype, cone_twist_joint_dtype, slider_joint_dtype]

assert body_dtype.itemsize == 140 and collider_dtype.itemsize == 64 and body_state_dtype.itemsize == 52
assert hinge_joint_dtype.itemsize == 104 and cone_twist_joint_dtype.itemsize == 120 and fixed_joint_dtype.itemsize == 48 and slider_joint_dtype.itemsize == 80

SPHERE, CAPSULE, CYLINDER, AABB, OBB, HULL = range(6)
OBJ_RIGID_BODY, OBJ_STATIC, OBJ_FORCE_FIELD, OBJ_TRIGGER = range(4)
NO_TRANSFORM = 0xFFFFFFFF
STEP_FAST, STEP_EXACT_ORDER, STEP_PERMUTED = 0, 1, 2


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("maxBodies", C.c_uint32), ("maxColliders", C.c_uint32), ("maxTransforms", C.c_uint32),
                ("maxPairs", C.c_uint32), ("maxManifolds", C.c_uint32), ("maxJointsPerKind", C.c_uint32), ("maxGridCells", C.c_uint32),
                ("cellSize", C.c_float), ("bigFactor", C.c_float), ("maxContacts", C.c_uint32), ("reserved", C.c_uint32 * 5)]


class ClothDesc(C.Structure):
    _fields_ = [("width", C.c_float), ("height", C.c_float), ("gridSizeX", C.c_uint32), ("gridSizeY", C.c_uint32), ("totalMass", C.c_float), ("stiffness", C.c_float),
                ("damping", C.c_float), ("gravityFactor", C.c_float)]


class HeightmapDesc(C.Structure):
    _fields_ = [("chunksPerDim", C.c_uint32), ("chunkSize", C.c_float), ("minCorner", C.c_float * 3), ("amplitudeScale", C.c_float), ("restitution", C.c_float), ("friction", C.c_float)]


class DistConfig(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("axis", C.c_int32), ("slabLo", C.c_float), ("slabHi", C.c_float), ("ghostWidth", C.c_float),
                ("maxGhostBodies", C.c_uint32), ("maxGhostColliders", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("numBodies", C.c_uint32), ("numColliders", C.c_uint32), ("numBroadphaseOverlaps", C.c_uint32), ("numCollisions", C.c_uint32),
                ("numContacts", C.c_uint32), ("numBigColliders", C.c_uint32), ("numGridCells", C.c_uint32), ("numColors", C.c_uint32 * 7),
                ("overflowPairs", C.c_uint32), ("overflowManifolds", C.c_uint32), ("overflowGrid", C.c_uint32), ("epaOverflows", C.c_uint32),
                ("gjkIterationCaps", C.c_uint32), ("assertCount", C.c_uint32), ("cellSize", C.c_float),
                ("msWorld", C.c_float), ("msBroad", C.c_float), ("msNarrow", C.c_float), ("msInit", C.c_float), ("msSolve", C.c_float),
                ("msIntegrate", C.c_float), ("msTotal", C.c_float), ("kernelLaunches", C.c_uint32), ("reserved", C.c_uint32 * 7)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


# every symbol include/b200phys.h declares (tests check that the library exports all of them)
EXPORTS = [
    "b200phys_version", "b200phys_error_string", "b200phys_create", "b200phys_destroy", "b200phys_set_stream", "b200phys_enable_timing", "b200phys_enable_graphs", "b200phys_upload_heightmap", "b200phys_env_create", "b200phys_env_set_seed", "b200phys_env_reset", "b200phys_env_step", "b200phys_cloth_create", "b200phys_cloth_count", "b200phys_cloth_set_fixed_vertices", "b200phys_cloth_set_properties",
    "b200phys_cloth_apply_wind", "b200phys_cloth_simulate", "b200phys_cloth_download",
    "b200phys_host_alloc", "b200phys_host_free",
    "b200phys_set_hulls", "b200phys_upload_bodies", "b200phys_upload_transforms", "b200phys_upload_colliders", "b200phys_upload_joints",
    "b200phys_upload_forces", "b200phys_upload_body_states", "b200phys_step", "b200phys_synchronize", "b200phys_download_bodies", "b200phys_get_stats", "b200phys_get_broadphase_reuse",
    "b200phys_download_aabbs", "b200phys_download_pairs", "b200phys_download_contacts", "b200phys_download_collisions",
    "b200phys_download_solve_order", "b200phys_get_timeline", "b200phys_dist_unique_id", "b200phys_create_distributed", "b200phys_step_distributed", "b200phys_dist_ghost_counts", "b200phys_halo_pack", "b200phys_halo_unpack", "b200phys_halo_clear", "b200phys_halo_record_bytes",
    "b200phys_halo_message_bytes", "b200phys_halo_pack_message", "b200phys_halo_unpack_message", "b200phys_migrate_pack", "b200phys_migrate_unpack", "b200phys_enable_events", "b200phys_download_events", "b200phys_upload_force_fields", "b200phys_set_global_force_field", "b200phys_download_trigger_events",
]

_lib = None


def dist_unique_id():
    """128-byte NCCL id from rank 0 (b200phys_dist_unique_id); hand it to every rank's Stepper(dist=...)."""
    lib = load_library()
    buf = (C.c_uint8 * 128)()
    _check(lib, lib.b200phys_dist_unique_id(buf), "b200phys_dist_unique_id")
    return bytes(buf)


def load_library():
    """Loads the CUDA library; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libb200phys.so is missing: build it with `make -C d3d12renderer_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u32, i32, f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float
    lib.b200phys_version.restype = i32
    lib.b200phys_error_string.restype = C.c_char_p; lib.b200phys_error_string.argtypes = [i32]
    lib.b200phys_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.b200phys_destroy.argtypes = [vp]
    lib.b200phys_set_stream.argtypes = [vp, vp]
    lib.b200phys_enable_timing.argtypes = [vp, i32]
    lib.b200phys_enable_graphs.argtypes = [vp, i32]
    lib.b200phys_upload_heightmap.argtypes = [vp, vp, vp]
    lib.b200phys_cloth_create.argtypes = [vp, vp, vp]; lib.b200phys_cloth_set_fixed_vertices.argtypes = [vp, C.c_uint32, vp, vp, vp, i32]
    lib.b200phys_cloth_set_properties.argtypes = [vp, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float]; lib.b200phys_cloth_apply_wind.argtypes = [vp, C.c_uint32, vp]
    lib.b200phys_cloth_simulate.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]; lib.b200phys_cloth_download.argtypes = [vp, C.c_uint32, vp, vp]
    lib.b200phys_host_alloc.restype = vp; lib.b200phys_host_alloc.argtypes = [C.c_size_t]
    lib.b200phys_host_free.restype = None; lib.b200phys_host_free.argtypes = [vp]
    lib.b200phys_set_hulls.argtypes = [vp, vp, vp, u32]
    lib.b200phys_upload_bodies.argtypes = [vp, vp, u32]
    lib.b200phys_upload_transforms.argtypes = [vp, vp, u32]
    lib.b200phys_upload_colliders.argtypes = [vp, vp, u32]
    lib.b200phys_upload_joints.argtypes = [vp, i32, vp, vp, u32]
    lib.b200phys_upload_forces.argtypes = [vp, vp, u32]
    lib.b200phys_upload_body_states.argtypes = [vp, vp, u32]
    lib.b200phys_step.argtypes = [vp, f32, u32, u32]
    lib.b200phys_synchronize.argtypes = [vp]
    lib.b200phys_download_bodies.argtypes = [vp, vp, u32]
    lib.b200phys_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.b200phys_get_broadphase_reuse.argtypes = [vp, C.POINTER(u32)]
    lib.b200phys_get_timeline.argtypes = [vp, vp, u32, vp, u32, C.POINTER(u32)]
    lib.b200phys_dist_unique_id.argtypes = [vp]
    lib.b200phys_create_distributed.argtypes = [C.POINTER(Config), C.POINTER(DistConfig), vp, C.POINTER(vp)]
    lib.b200phys_step_distributed.argtypes = [vp, f32, u32, u32]
    lib.b200phys_dist_ghost_counts.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    lib.b200phys_download_aabbs.argtypes = [vp, vp, C.POINTER(u32)]
    lib.b200phys_download_pairs.argtypes = [vp, vp, C.POINTER(u32)]
    lib.b200phys_download_contacts.argtypes = [vp, vp, vp, C.POINTER(u32)]
    lib.b200phys_download_collisions.argtypes = [vp, vp, vp, C.POINTER(u32)]
    lib.b200phys_download_solve_order.argtypes = [vp, i32, vp, C.POINTER(u32)]
    lib.b200phys_halo_pack.argtypes = [vp, i32, f32, f32, vp, vp, u32, u32, C.POINTER(u32)]
    lib.b200phys_halo_unpack.argtypes = [vp, vp, u32, vp, u32]
    lib.b200phys_halo_clear.argtypes = [vp]
    lib.b200phys_halo_message_bytes.restype = C.c_uint64; lib.b200phys_halo_message_bytes.argtypes = [u32, u32]
    lib.b200phys_halo_pack_message.argtypes = [vp, i32, f32, f32, vp, u32, u32]
    lib.b200phys_halo_unpack_message.argtypes = [vp, vp, u32, u32]
    lib.b200phys_migrate_pack.argtypes = [vp, i32, f32, f32, vp, vp, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.b200phys_migrate_unpack.argtypes = [vp, vp, u32, u32, C.POINTER(u32)]
    lib.b200phys_enable_events.argtypes = [vp, i32]
    lib.b200phys_upload_force_fields.argtypes = [vp, vp, u32]
    lib.b200phys_set_global_force_field.argtypes = [vp, vp]
    lib.b200phys_download_trigger_events.argtypes = [vp, vp, u32, C.POINTER(u32), vp, u32, C.POINTER(u32)]
    lib.b200phys_download_events.argtypes = [vp, vp, u32, C.POINTER(u32), vp, u32, C.POINTER(u32)]
    lib.b200phys_halo_record_bytes.restype = u32; lib.b200phys_halo_record_bytes.argtypes = [i32]
    _lib = lib
    return lib


class B200PhysError(RuntimeError):
    pass


def _check(lib, code, what):
    if code != 0:
        raise B200PhysError("%s failed: %d (%s)" % (what, code, lib.b200phys_error_string(code).decode()))


def _ptr(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


class PinnedArray:
    """numpy view over page-locked memory from b200phys_host_alloc (direct DMA uploads/downloads)."""

    def __init__(self, dtype, count):
        self.lib = load_library()
        self.nbytes = max(1, np.dtype(dtype).itemsize * count)
        self.ptr = self.lib.b200phys_host_alloc(self.nbytes)
        if not self.ptr:
            raise B200PhysError("b200phys_host_alloc failed")
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=count)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.b200phys_host_free(self.ptr)
            self.ptr = None


class Stepper:
    """Thin object wrapper over one b200phys_context (one per GPU / rank)."""

    def __init__(self, max_bodies, max_colliders, max_transforms=16, device=0, max_pairs=0, max_manifolds=0, max_contacts=0,
                 max_joints=0, max_grid_cells=0, cell_size=0.0, big_factor=0.0, dist=None, bp_reuse=True):
        self.lib = load_library()
        cfg = Config()
        cfg.device = device; cfg.maxBodies = max_bodies; cfg.maxColliders = max_colliders; cfg.maxTransforms = max_transforms
        cfg.maxPairs = max_pairs; cfg.maxManifolds = max_manifolds; cfg.maxContacts = max_contacts; cfg.maxJointsPerKind = max_joints
        cfg.maxGridCells = max_grid_cells; cfg.cellSize = cell_size; cfg.bigFactor = big_factor
        if not bp_reuse:
            cfg.reserved[2] = 1          # the grid pipeline runs in every step (no candidate reuse)
        self.ctx = C.c_void_p()
        if dist is None:
            _check(self.lib, self.lib.b200phys_create(C.byref(cfg), C.byref(self.ctx)), "b200phys_create")
        else:       # dist = (DistConfig, 128-byte NCCL id): the library owns the NCCL communicator and the halo exchange (k_dist.cu)
            dc, uid = dist
            buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
            _check(self.lib, self.lib.b200phys_create_distributed(C.byref(cfg), C.byref(dc), buf, C.byref(self.ctx)), "b200phys_create_distributed")
        self.num_bodies = 0

    def close(self):
        if self.ctx:
            self.lib.b200phys_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(self.lib, self.lib.b200phys_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def enable_timing(self, on=True):
        _check(self.lib, self.lib.b200phys_enable_timing(self.ctx, 1 if on else 0), "enable_timing")

    # ---- cloth
    def cloth_create(self, width, height, gx, gy, total_mass, stiffness=0.5, damping=0.3, gravity_factor=1.0):
        d = ClothDesc(width, height, gx, gy, total_mass, stiffness, damping, gravity_factor); out = C.c_uint32(0)
        _check(self.lib, self.lib.b200phys_cloth_create(self.ctx, C.byref(d), C.byref(out)), "cloth_create")
        self._cloth_sizes = getattr(self, "_cloth_sizes", {}); self._cloth_sizes[out.value] = gx * gy
        return out.value

    def cloth_set_fixed(self, cloth, rotation, position, scale=(1, 1, 1), move_rigid=False):
        r = (C.c_float * 4)(*[float(x) for x in rotation]); p = (C.c_float * 3)(*[float(x) for x in position]); s = (C.c_float * 3)(*[float(x) for x in scale])
        _check(self.lib, self.lib.b200phys_cloth_set_fixed_vertices(self.ctx, cloth, r, p, s, 1 if move_rigid else 0), "cloth_set_fixed_vertices")

    def cloth_set_properties(self, cloth, total_mass, stiffness, damping, gravity_factor):
        _check(self.lib, self.lib.b200phys_cloth_set_properties(self.ctx, cloth, C.c_float(total_mass), C.c_float(stiffness), C.c_float(damping), C.c_float(gravity_factor)), "cloth_set_properties")

    def cloth_wind(self, force, cloth=0xFFFFFFFF):
        f = (C.c_float * 3)(*[float(x) for x in force])
        _check(self.lib, self.lib.b200phys_cloth_apply_wind(self.ctx, cloth, f), "cloth_apply_wind")

    def cloth_simulate(self, vel_iters, pos_iters, drift_iters, dt):
        _check(self.lib, self.lib.b200phys_cloth_simulate(self.ctx, vel_iters, pos_iters, drift_iters, C.c_float(dt)), "cloth_simulate")

    def cloth_state(self, cloth):
        n = self._cloth_sizes[cloth]
        p = np.zeros((n, 3), dtype=np.float32); v = np.zeros((n, 3), dtype=np.float32)
        _check(self.lib, self.lib.b200phys_cloth_download(self.ctx, cloth, _ptr(p), _ptr(v)), "cloth_download"); return p, v

    def upload_heightmap(self, heightmap):
        """heightmap: dict(chunks_per_dim, chunk_size, min_corner, amplitude_scale, restitution, friction, heights uint16 [chunks, 129, 129]) or None."""
        if heightmap is None:
            _check(self.lib, self.lib.b200phys_upload_heightmap(self.ctx, None, None), "upload_heightmap"); return
        d = HeightmapDesc(); d.chunksPerDim = int(heightmap["chunks_per_dim"]); d.chunkSize = float(heightmap["chunk_size"])
        for k in range(3):
            d.minCorner[k] = float(heightmap["min_corner"][k])
        d.amplitudeScale = float(heightmap["amplitude_scale"]); d.restitution = float(heightmap["restitution"]); d.friction = float(heightmap["friction"])
        h = np.ascontiguousarray(heightmap["heights"], dtype=np.uint16)
        assert h.size == d.chunksPerDim * d.chunksPerDim * 129 * 129
        _check(self.lib, self.lib.b200phys_upload_heightmap(self.ctx, C.byref(d), _ptr(h)), "upload_heightmap")

    def enable_graphs(self, on=True):
        _check(self.lib, self.lib.b200phys_enable_graphs(self.ctx, 1 if on else 0), "enable_graphs")

    def set_hulls(self, vertices, offsets):
        v = np.ascontiguousarray(vertices, dtype=np.float32); o = np.ascontiguousarray(offsets, dtype=np.uint32)
        _check(self.lib, self.lib.b200phys_set_hulls(self.ctx, _ptr(v), _ptr(o), len(o) - 1), "set_hulls")

    def upload_bodies(self, bodies):
        assert bodies.dtype == body_dtype and bodies.flags.c_contiguous
        _check(self.lib, self.lib.b200phys_upload_bodies(self.ctx, _ptr(bodies), len(bodies)), "upload_bodies")
        self.num_bodies = len(bodies)

    def upload_transforms(self, transforms):
        assert transforms.dtype == transform_dtype and transforms.flags.c_contiguous
        _check(self.lib, self.lib.b200phys_upload_transforms(self.ctx, _ptr(transforms), len(transforms)), "upload_transforms")

    def upload_colliders(self, colliders):
        assert colliders.dtype == collider_dtype and colliders.flags.c_contiguous
        _check(self.lib, self.lib.b200phys_upload_colliders(self.ctx, _ptr(colliders), len(colliders)), "upload_colliders")

    def upload_joints(self, kind, joints, pairs):
        assert joints.dtype == JOINT_DTYPES[kind] and pairs.dtype == body_pair_dtype
        _check(self.lib, self.lib.b200phys_upload_joints(self.ctx, kind, _ptr(np.ascontiguousarray(joints)), _ptr(np.ascontiguousarray(pairs)), len(joints)), "upload_joints")

    def upload_body_states(self, states):
        assert states.dtype == body_state_dtype and states.flags.c_contiguous
        _check(self.lib, self.lib.b200phys_upload_body_states(self.ctx, _ptr(states), len(states)), "upload_body_states")

    def upload_forces(self, force_torque):
        ft = np.ascontiguousarray(force_torque, dtype=np.float32)
        _check(self.lib, self.lib.b200phys_upload_forces(self.ctx, _ptr(ft), ft.shape[0]), "upload_forces")

    def step(self, dt, iterations, flags=STEP_FAST):
        _check(self.lib, self.lib.b200phys_step(self.ctx, dt, iterations, flags), "step")

    def step_distributed(self, dt, iterations, flags=STEP_FAST):
        """Ghost exchange with the slab neighbours (NCCL inside the library) + the step."""
        _check(self.lib, self.lib.b200phys_step_distributed(self.ctx, dt, iterations, flags), "step_distributed")

    def dist_ghost_counts(self):
        b, c = C.c_uint32(0), C.c_uint32(0)
        _check(self.lib, self.lib.b200phys_dist_ghost_counts(self.ctx, C.byref(b), C.byref(c)), "dist_ghost_counts")
        return b.value, c.value

    def synchronize(self):
        _check(self.lib, self.lib.b200phys_synchronize(self.ctx), "synchronize")

    def download_bodies(self, out=None):
        if out is None:
            out = np.empty(self.num_bodies, dtype=body_state_dtype)
        _check(self.lib, self.lib.b200phys_download_bodies(self.ctx, _ptr(out), len(out)), "download_bodies")
        return out

    def stats(self, check=True):
        st = Stats()
        code = self.lib.b200phys_get_stats(self.ctx, C.byref(st))
        if check:
            _check(self.lib, code, "get_stats")
        return st

    def broadphase_reuse(self):
        """Candidate reuse of the broad phase: dict(steps, rebuilds, candidates, pairs, margin, age) (b200phys_get_broadphase_reuse)."""
        out = (C.c_uint32 * 8)()
        _check(self.lib, self.lib.b200phys_get_broadphase_reuse(self.ctx, out), "get_broadphase_reuse")
        f = np.frombuffer(np.array(list(out), dtype=np.uint32).tobytes(), dtype=np.float32)
        return dict(steps=int(out[0]), rebuilds=int(out[1]), candidates=int(out[2]), pairs=int(out[3]), margin=float(f[4]), age=int(out[5]),
                    moved_per_step=float(f[6]), mean_extent=float(f[7]))

    def timeline(self):
        """[(kernel group name, ms)] of the last step (enable_timing(True) first)."""
        names = C.create_string_buffer(4096); ms = (C.c_float * 64)(); n = C.c_uint32(0)
        _check(self.lib, self.lib.b200phys_get_timeline(self.ctx, names, 4096, ms, 64, C.byref(n)), "get_timeline")
        labels = names.value.decode().split("\n")[:n.value]
        return [(labels[i], float(ms[i])) for i in range(min(n.value, len(labels)))]

    def _count_then_fetch(self, fn, dtypes, name):
        n = C.c_uint32(0)
        nulls = [None] * len(dtypes)
        _check(self.lib, fn(self.ctx, *nulls, C.byref(n)), name)
        outs = [np.empty(n.value, dtype=dt) for dt in dtypes]
        if n.value:
            _check(self.lib, fn(self.ctx, *[_ptr(o) for o in outs], C.byref(n)), name)
        return outs

    def download_aabbs(self):
        n = C.c_uint32(0)
        _check(self.lib, self.lib.b200phys_download_aabbs(self.ctx, None, C.byref(n)), "download_aabbs")
        out = np.empty((n.value, 6), dtype=np.float32)
        if n.value:
            _check(self.lib, self.lib.b200phys_download_aabbs(self.ctx, _ptr(out), C.byref(n)), "download_aabbs")
        return out

    def download_pairs(self):
        return self._count_then_fetch(self.lib.b200phys_download_pairs, [pair_dtype], "download_pairs")[0]

    def download_contacts(self):
        return self._count_then_fetch(self.lib.b200phys_download_contacts, [contact_dtype, body_pair_dtype], "download_contacts")

    def download_collisions(self):
        return self._count_then_fetch(self.lib.b200phys_download_collisions, [pair_dtype, np.dtype("u1")], "download_collisions")

    # ---- multi-GPU halo (device pointers: e.g. torch tensors' data_ptr())
    def halo_pack(self, axis, lo, hi, dst_bodies_ptr, dst_colliders_ptr, cap_bodies, cap_colliders):
        counts = (C.c_uint32 * 2)()
        _check(self.lib, self.lib.b200phys_halo_pack(self.ctx, axis, lo, hi, C.c_void_p(dst_bodies_ptr), C.c_void_p(dst_colliders_ptr), cap_bodies, cap_colliders, counts), "halo_pack")
        return counts[0], counts[1]

    def halo_unpack(self, src_bodies_ptr, num_bodies, src_colliders_ptr, num_colliders):
        _check(self.lib, self.lib.b200phys_halo_unpack(self.ctx, C.c_void_p(src_bodies_ptr), num_bodies, C.c_void_p(src_colliders_ptr), num_colliders), "halo_unpack")

    # ---- force fields and triggers
    def upload_force_fields(self, forces, global_force=None):
        f = np.ascontiguousarray(forces, dtype=np.float32).reshape(-1, 3)
        _check(self.lib, self.lib.b200phys_upload_force_fields(self.ctx, _ptr(f) if len(f) else None, len(f)), "upload_force_fields")
        g = np.zeros(3, dtype=np.float32) if global_force is None else np.ascontiguousarray(global_force, dtype=np.float32)
        _check(self.lib, self.lib.b200phys_set_global_force_field(self.ctx, _ptr(g)), "set_global_force_field")

    def download_trigger_events(self):
        """-> (enter [n,2], leave [m,2]) as (triggerIndex, rigidBodyIndex); call once after every step."""
        ne, nl = C.c_uint32(0), C.c_uint32(0)
        cap = 1 << 16
        e = np.zeros((cap, 2), dtype=np.uint32); l = np.zeros((cap, 2), dtype=np.uint32)
        _check(self.lib, self.lib.b200phys_download_trigger_events(self.ctx, _ptr(e), cap, C.byref(ne), _ptr(l), cap, C.byref(nl)), "download_trigger_events")
        return e[:ne.value].copy(), l[:nl.value].copy()

    # ---- collision begin / end events
    def enable_events(self, on=True):
        _check(self.lib, self.lib.b200phys_enable_events(self.ctx, 1 if on else 0), "enable_events")

    def download_events(self):
        """-> (begin events [begin_event_dtype], end events [pair_dtype]) of the last step."""
        nb, ne = C.c_uint32(0), C.c_uint32(0)
        self.lib.b200phys_download_events(self.ctx, None, 0, C.byref(nb), None, 0, C.byref(ne))
        b = np.zeros(nb.value, dtype=begin_event_dtype); e = np.zeros(ne.value, dtype=pair_dtype)
        _check(self.lib, self.lib.b200phys_download_events(self.ctx, _ptr(b) if nb.value else None, nb.value, C.byref(nb), _ptr(e) if ne.value else None, ne.value, C.byref(ne)), "download_events")
        return b, e

    def halo_message_bytes(self, cap_bodies, cap_colliders):
        return int(self.lib.b200phys_halo_message_bytes(cap_bodies, cap_colliders))

    def halo_pack_message(self, axis, lo, hi, dst_ptr, cap_bodies, cap_colliders):
        """Asynchronous: fills the fixed-size message at device pointer dst_ptr on the context's stream."""
        _check(self.lib, self.lib.b200phys_halo_pack_message(self.ctx, axis, lo, hi, C.c_void_p(dst_ptr), cap_bodies, cap_colliders), "halo_pack_message")

    def halo_unpack_message(self, src_ptr, cap_bodies, cap_colliders):
        _check(self.lib, self.lib.b200phys_halo_unpack_message(self.ctx, C.c_void_p(src_ptr), cap_bodies, cap_colliders), "halo_unpack_message")

    def migrate_pack(self, axis, lo, hi, dst_down_ptr, dst_up_ptr, cap_bodies, cap_colliders, num_owned=None):
        """-> (dest[num_owned] uint32: 0 stays / 1 down / 2 up, counts (bodies down, colliders down, bodies up, colliders up))."""
        num_owned = self.num_bodies if num_owned is None else num_owned
        dest = np.zeros(max(1, num_owned), dtype=np.uint32); counts = (C.c_uint32 * 4)()
        _check(self.lib, self.lib.b200phys_migrate_pack(self.ctx, axis, lo, hi, C.c_void_p(dst_down_ptr), C.c_void_p(dst_up_ptr), cap_bodies, cap_colliders,
                                                        dest.ctypes.data_as(C.POINTER(C.c_uint32)), counts), "migrate_pack")
        self.num_bodies = num_owned - int(counts[0]) - int(counts[2])
        return dest[:num_owned], tuple(int(x) for x in counts)

    def migrate_unpack(self, src_ptr, cap_bodies, cap_colliders):
        counts = (C.c_uint32 * 2)()
        _check(self.lib, self.lib.b200phys_migrate_unpack(self.ctx, C.c_void_p(src_ptr), cap_bodies, cap_colliders, counts), "migrate_unpack")
        self.num_bodies += int(counts[0])
        return int(counts[0]), int(counts[1])

    def halo_clear(self):
        _check(self.lib, self.lib.b200phys_halo_clear(self.ctx), "halo_clear")

    def download_solve_order(self, kind):
        n = C.c_uint32(0)
        _check(self.lib, self.lib.b200phys_download_solve_order(self.ctx, kind, None, C.byref(n)), "download_solve_order")
        out = np.empty(n.value, dtype=np.uint32)
        if n.value:
            _check(self.lib, self.lib.b200phys_download_solve_order(self.ctx, kind, _ptr(out), C.byref(n)), "download_solve_order")
        return out
