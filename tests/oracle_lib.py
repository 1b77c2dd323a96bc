"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product path)."""
import ctypes as C
import os
import subprocess
import numpy as np

from d3d12renderer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    path = LIB
    if os.environ.get("B200_ORACLE_VARIANT") == "fast":      # bench.py only: the AVX2 / fast-math build as the CPU timing baseline
        fast = os.path.join(ORACLE_DIR, "liboracle_fast.so")
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            flags = ""
        if os.path.exists(fast) and " avx2" in flags and " fma" in flags:      # never risk an illegal instruction inside the bench process
            path = fast
    try:
        lib = C.CDLL(path)
    except OSError:
        lib = C.CDLL(LIB)
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    lib.oracle_create.restype = vp
    for name in ["oracle_destroy", "oracle_stage_detect", "oracle_stage_world_and_aabbs"]:
        getattr(lib, name).argtypes = [vp]
    lib.oracle_set_pair_order.argtypes = [vp, C.c_int]
    lib.oracle_set_hulls.argtypes = [vp, vp, vp, u32]
    lib.oracle_set_hull_triangles.argtypes = [vp, u32, vp, u32]
    lib.oracle_set_bodies.argtypes = [vp, vp, u32]
    lib.oracle_set_transforms.argtypes = [vp, vp, u32]
    lib.oracle_set_colliders.argtypes = [vp, vp, u32]
    lib.oracle_set_joints.argtypes = [vp, C.c_int, vp, vp, u32]
    lib.oracle_step.argtypes = [vp, f32, u32] + [vp] * 7
    lib.oracle_stage_forces.argtypes = [vp, f32]
    lib.oracle_stage_init.argtypes = [vp, f32]
    lib.oracle_stage_solve.argtypes = [vp, u32] + [vp] * 7
    lib.oracle_stage_integrate.argtypes = [vp, f32]
    lib.oracle_set_contacts.argtypes = [vp, vp, vp, u32]
    for name in ["oracle_num_bodies", "oracle_num_colliders", "oracle_num_pairs", "oracle_num_contacts", "oracle_num_collisions"]:
        getattr(lib, name).argtypes = [vp]; getattr(lib, name).restype = u32
    lib.oracle_get_bodies.argtypes = [vp, vp]
    lib.oracle_get_body_states.argtypes = [vp, vp]
    lib.oracle_get_global_velocities.argtypes = [vp, vp]
    lib.oracle_get_aabbs.argtypes = [vp, vp]
    lib.oracle_get_world_colliders.argtypes = [vp, vp, vp]
    lib.oracle_get_pairs.argtypes = [vp, vp]
    lib.oracle_brute_pairs.argtypes = [vp, vp, u32]; lib.oracle_brute_pairs.restype = u32
    lib.oracle_get_contacts.argtypes = [vp, vp, vp]
    lib.oracle_get_collisions.argtypes = [vp, vp, vp]
    lib.oracle_get_contact_impulses.argtypes = [vp, vp]
    lib.oracle_get_narrow_stats.argtypes = [vp, vp]
    lib.oracle_get_stage_ms.argtypes = [vp, vp]
    lib.oracle_collide_pair.argtypes = [vp, vp, vp, vp, vp]; lib.oracle_collide_pair.restype = C.c_int
    lib.oracle_mass_properties.argtypes = [vp, vp, vp, vp, vp]
    lib.oracle_cloth_create.argtypes = [f32, f32, u32, u32, f32, f32, f32, f32]; lib.oracle_cloth_create.restype = vp
    lib.oracle_cloth_destroy.argtypes = [vp]; lib.oracle_cloth_set_fixed.argtypes = [vp, vp, vp, vp, C.c_int]; lib.oracle_cloth_set_properties.argtypes = [vp, f32, f32, f32, f32]
    lib.oracle_cloth_wind.argtypes = [vp, vp]; lib.oracle_cloth_simulate.argtypes = [vp, u32, u32, u32, f32]; lib.oracle_cloth_get.argtypes = [vp, vp, vp]
    lib.oracle_set_heightmap.argtypes = [vp, u32, f32, vp, f32, f32, f32, vp]
    lib.oracle_heightmap_height_at.argtypes = [vp, f32, f32]; lib.oracle_heightmap_height_at.restype = f32
    lib.oracle_combine_mass.argtypes = [vp, vp, vp, u32, vp, vp, vp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, scene=None, canonical=True):
        self.lib = load()
        self.ctx = C.c_void_p(self.lib.oracle_create())
        self.lib.oracle_set_pair_order(self.ctx, 1 if canonical else 0)
        if scene is not None:
            self.set_scene(scene)

    def __del__(self):
        try:
            if self.ctx:
                self.lib.oracle_destroy(self.ctx); self.ctx = None
        except Exception:
            pass

    def set_scene(self, scene):
        if "hull_vertices" in scene:
            v = np.ascontiguousarray(scene["hull_vertices"], dtype=np.float32); o = np.ascontiguousarray(scene["hull_offsets"], dtype=np.uint32)
            self.lib.oracle_set_hulls(self.ctx, _p(v), _p(o), len(o) - 1)
            for h, tris in enumerate(scene.get("hull_triangles", [])):
                t = np.ascontiguousarray(tris, dtype=np.uint32)
                self.lib.oracle_set_hull_triangles(self.ctx, h, _p(t), len(t) // 3)
        b = np.ascontiguousarray(scene["bodies"]); t = np.ascontiguousarray(scene["transforms"]); c = np.ascontiguousarray(scene["colliders"])
        self.lib.oracle_set_bodies(self.ctx, _p(b), len(b))
        self.lib.oracle_set_transforms(self.ctx, _p(t), len(t))
        self.lib.oracle_set_colliders(self.ctx, _p(c), len(c))
        for kind, (joints, pairs) in scene.get("joints", {}).items():
            j = np.ascontiguousarray(joints); p = np.ascontiguousarray(pairs)
            self.lib.oracle_set_joints(self.ctx, kind, _p(j), _p(p), len(j))
        if "heightmap" in scene:
            h = scene["heightmap"]; mc = np.ascontiguousarray(h["min_corner"], dtype=np.float32); hh = np.ascontiguousarray(h["heights"], dtype=np.uint16)
            self.lib.oracle_set_heightmap(self.ctx, h["chunks_per_dim"], h["chunk_size"], _p(mc), h["amplitude_scale"], h["restitution"], h["friction"], _p(hh))

    def heightmap_height_at(self, x, z): return float(self.lib.oracle_heightmap_height_at(self.ctx, x, z))

    def set_bodies(self, bodies):
        b = np.ascontiguousarray(bodies)
        self.lib.oracle_set_bodies(self.ctx, _p(b), len(b))

    def step(self, dt, iterations, contact_order=None, joint_orders=None):
        jo = [None] * 6
        if joint_orders:
            for k, o in joint_orders.items():
                jo[k] = np.ascontiguousarray(o, dtype=np.uint32)
        co = np.ascontiguousarray(contact_order, dtype=np.uint32) if contact_order is not None else None
        self.lib.oracle_step(self.ctx, dt, iterations, _p(co), *[_p(x) for x in jo])

    # staged
    def detect(self): self.lib.oracle_stage_detect(self.ctx)
    def world(self): self.lib.oracle_stage_world_and_aabbs(self.ctx)
    def forces(self, dt): self.lib.oracle_stage_forces(self.ctx, dt)
    def init(self, dt): self.lib.oracle_stage_init(self.ctx, dt)

    def solve(self, iterations, contact_order=None, joint_orders=None):
        jo = [None] * 6
        if joint_orders:
            for k, o in joint_orders.items():
                jo[k] = np.ascontiguousarray(o, dtype=np.uint32)
        co = np.ascontiguousarray(contact_order, dtype=np.uint32) if contact_order is not None else None
        self.lib.oracle_stage_solve(self.ctx, iterations, _p(co), *[_p(x) for x in jo])

    def integrate(self, dt): self.lib.oracle_stage_integrate(self.ctx, dt)

    def body_states(self):
        out = np.empty(self.lib.oracle_num_bodies(self.ctx), dtype=abi.body_state_dtype)
        self.lib.oracle_get_body_states(self.ctx, _p(out)); return out

    def bodies(self):
        out = np.empty(self.lib.oracle_num_bodies(self.ctx), dtype=abi.body_dtype)
        self.lib.oracle_get_bodies(self.ctx, _p(out)); return out

    def aabbs(self):
        out = np.empty((self.lib.oracle_num_colliders(self.ctx), 6), dtype=np.float32)
        self.lib.oracle_get_aabbs(self.ctx, _p(out)); return out

    def pairs(self):
        out = np.empty(self.lib.oracle_num_pairs(self.ctx), dtype=abi.pair_dtype)
        self.lib.oracle_get_pairs(self.ctx, _p(out)); return out

    def brute_pairs(self):
        n = self.lib.oracle_brute_pairs(self.ctx, None, 0)
        out = np.empty(n, dtype=abi.pair_dtype)
        self.lib.oracle_brute_pairs(self.ctx, _p(out), n); return out

    def contacts(self):
        n = self.lib.oracle_num_contacts(self.ctx)
        c = np.empty(n, dtype=abi.contact_dtype); b = np.empty(n, dtype=abi.body_pair_dtype)
        self.lib.oracle_get_contacts(self.ctx, _p(c), _p(b)); return c, b

    def collisions(self):
        n = self.lib.oracle_num_collisions(self.ctx)
        p = np.empty(n, dtype=abi.pair_dtype); k = np.empty(n, dtype=np.uint8)
        self.lib.oracle_get_collisions(self.ctx, _p(p), _p(k)); return p, k

    def set_force_fields(self, forces, global_force=None):
        f = np.ascontiguousarray(forces, dtype=np.float32).reshape(-1, 3)
        g = np.zeros(3, dtype=np.float32) if global_force is None else np.ascontiguousarray(global_force, dtype=np.float32)
        self.lib.oracle_set_force_fields(self.ctx, _p(f) if len(f) else None, C.c_uint32(len(f)), _p(g))

    def trigger_events(self):
        """(enter [n,2], leave [m,2]) as (triggerIndex, rigidBodyIndex) of the last step."""
        self.lib.oracle_num_trigger_events.restype = C.c_uint32
        nl = C.c_uint32(0); ne = self.lib.oracle_num_trigger_events(self.ctx, C.byref(nl))
        e = np.zeros((ne, 2), dtype=np.uint32); l = np.zeros((nl.value, 2), dtype=np.uint32)
        self.lib.oracle_get_trigger_events(self.ctx, _p(e), _p(l)); return e, l

    def num_interactions(self):
        self.lib.oracle_num_interactions.restype = C.c_uint32
        return self.lib.oracle_num_interactions(self.ctx)

    def events(self):
        """Collision begin / end events of the last step: (begin pairs [n,2], begin payload [n,9] = point, normal, relative velocity, end pairs [m,2])."""
        self.lib.oracle_num_begin_events.restype = C.c_uint32; self.lib.oracle_num_end_events.restype = C.c_uint32
        nb, ne = self.lib.oracle_num_begin_events(self.ctx), self.lib.oracle_num_end_events(self.ctx)
        bp = np.zeros((nb, 2), dtype=np.uint32); pl = np.zeros((nb, 9), dtype=np.float32); ep = np.zeros((ne, 2), dtype=np.uint32)
        self.lib.oracle_get_events(self.ctx, _p(bp), _p(pl), _p(ep)); return bp, pl, ep

    def narrow_stats(self):
        out = np.zeros(8, dtype=np.uint32); self.lib.oracle_get_narrow_stats(self.ctx, _p(out))
        return dict(zip(["gjkIterCap", "gjkErrors", "epaNotConverged", "epaOutOfMemory", "asserts", "epaMaxTriangles", "epaMaxEdges", "epaMaxPoints"], out.tolist()))

    def stage_ms(self):
        out = np.zeros(7, dtype=np.float64); self.lib.oracle_get_stage_ms(self.ctx, _p(out)); return out

    def collide_pair(self, col_a, col_b):
        a = np.ascontiguousarray(col_a); b = np.ascontiguousarray(col_b)
        normal = np.zeros(3, dtype=np.float32); pd = np.zeros((4, 4), dtype=np.float32)
        n = self.lib.oracle_collide_pair(self.ctx, _p(a), _p(b), _p(normal), _p(pd))
        return n, normal, pd[:n]

    def overlap_pair(self, col_a, col_b):
        a = np.ascontiguousarray(col_a); b = np.ascontiguousarray(col_b)
        self.lib.oracle_overlap_pair.restype = C.c_int
        return bool(self.lib.oracle_overlap_pair(self.ctx, _p(a), _p(b)))

    def mass_properties(self, collider):
        c = np.ascontiguousarray(collider)
        inertia = np.zeros(9, dtype=np.float32); cog = np.zeros(3, dtype=np.float32); mass = np.zeros(1, dtype=np.float32)
        self.lib.oracle_mass_properties(self.ctx, _p(c), _p(inertia), _p(cog), _p(mass))
        return inertia, cog, float(mass[0])
