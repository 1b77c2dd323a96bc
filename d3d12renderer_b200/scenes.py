"""Synthetic scene generators for the BASELINE.json configurations.

Host-side set-up code (not on the per-step hot path): produces the flat boundary arrays (abi.body_dtype, abi.collider_dtype,
abi.transform_dtype, joints) that both the CUDA stepper and the parity oracle consume. All randomness comes from the
reference's xorshift64 generator (src/core/random.h:5-50) so scenes are reproducible from their seed alone.
Mass properties follow collider_union::calculatePhysicsProperties (src/physics/physics.cpp:1416-1588) in float32.
"""
import numpy as np
from . import abi

F = np.float32
M_PI = F(3.14159265359)  # reference src/core/math.h:13


class Xorshift64:
    """random_number_generator (reference src/core/random.h:5-50)."""

    def __init__(self, seed):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def u64(self):
        x = self.state
        x ^= (x << 13) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 7
        x ^= (x << 17) & 0xFFFFFFFFFFFFFFFF
        self.state = x
        return x

    def f01(self):
        return F(F(self.u64() & 0xFFFFFFFF) / F(4294967295.0))

    def between(self, lo, hi):
        t = F((self.f01() - F(0)) / (F(1) - F(0)))
        return F(F(lo) + t * (F(hi) - F(lo)))


def xorshift_floats01(seed, count):
    """count values of randomFloat01() as a float32 array (vectorised output, sequential generator)."""
    out = np.empty(count, dtype=np.float64)
    x = seed & 0xFFFFFFFFFFFFFFFF
    mask = 0xFFFFFFFFFFFFFFFF
    for i in range(count):
        x ^= (x << 13) & mask
        x ^= x >> 7
        x ^= (x << 17) & mask
        out[i] = x & 0xFFFFFFFF
    return (out.astype(np.float32) / F(4294967295.0)).astype(np.float32)


def sphere_mass(radius, density):
    r = F(radius)
    sq = F(r * r); sq_pi = F(M_PI * sq)
    volume = F(F(F(4.0) / F(3.0)) * sq_pi * r)
    mass = F(volume * F(density))
    inertia = F(F(F(F(2.0) / F(5.0)) * mass) * r * r)
    return mass, inertia


def new_bodies(n):
    b = np.zeros(n, dtype=abi.body_dtype)
    b["rotation"][:, 3] = 1.0
    b["gravityFactor"] = 1.0
    b["linearDamping"] = 0.4   # rigid_body.h:21 defaults
    b["angularDamping"] = 0.4
    return b


def set_sphere_bodies(bodies, idx, radius, density):
    mass, inertia = sphere_mass(radius, density)
    bodies["invMass"][idx] = F(1.0) / mass
    inv_i = F(1.0) / inertia
    for k in (0, 4, 8):
        bodies["invInertia"][idx, k] = inv_i


def static_box(colliders, i, center, radius, transform_index, restitution=0.1, friction=1.0, density=4.0):
    c = np.asarray(center, dtype=np.float32); r = np.asarray(radius, dtype=np.float32)
    colliders["shape"][i, 0:3] = c - r
    colliders["shape"][i, 3:6] = c + r
    colliders["restitution"][i] = restitution; colliders["friction"][i] = friction; colliders["density"][i] = density
    colliders["type"][i] = abi.AABB; colliders["objectType"][i] = abi.OBJ_STATIC
    colliders["objectIndex"][i] = 0; colliders["parentTransform"][i] = transform_index


def spheres_on_ground(n_side=32, spacing=1.25, mode="drop", seed=15681923, radius=0.5):
    """BASELINE config 1: n_side^2 dynamic spheres over the reference's ground box (application.cpp:215,248-251).
    mode 'drop': y = 0.5 + rand[0,10] (free fall, few contacts); mode 'rest': y = 0.45 (every sphere penetrates 0.05)."""
    n = n_side * n_side
    bodies = new_bodies(n)
    ix, iz = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    half = F((n_side - 1) * spacing * 0.5)
    bodies["position"][:, 0] = (ix.ravel().astype(np.float32) * F(spacing) - half)
    bodies["position"][:, 2] = (iz.ravel().astype(np.float32) * F(spacing) - half)
    if mode == "drop":
        r = xorshift_floats01(seed, n)
        bodies["position"][:, 1] = F(0.5) + r * F(10.0)
    else:
        bodies["position"][:, 1] = F(0.45)
    set_sphere_bodies(bodies, slice(None), radius, 1.0)
    colliders = np.zeros(n + 1, dtype=abi.collider_dtype)
    colliders["shape"][:n, 3] = radius
    colliders["restitution"][:n] = 0.1; colliders["friction"][:n] = 0.5; colliders["density"][:n] = 1.0
    colliders["type"][:n] = abi.SPHERE; colliders["objectType"][:n] = abi.OBJ_RIGID_BODY
    colliders["objectIndex"][:n] = np.arange(n, dtype=np.uint32); colliders["parentTransform"][:n] = abi.NO_TRANSFORM
    transforms = np.zeros(1, dtype=abi.transform_dtype)
    transforms["rotation"][0, 3] = 1.0; transforms["position"][0] = (0.0, -4.0, 0.0)
    extent = max(30.0, float(half) + 2.0)
    static_box(colliders, n, (0, 0, 0), (extent, 4.0, extent), 0)
    return {"bodies": bodies, "colliders": colliders, "transforms": transforms, "name": "spheres_on_ground_%d_%s" % (n, mode)}


def sphere_box_drop(nx=64, ny=64, nz=64, spacing=1.02, jitter=0.005, seed=519431, radius=0.5, base_y=0.6, walls=True):
    """BASELINE config 2: nx*ny*nz spheres on a jittered lattice inside an open box of 5 static AABBs (floor + 4 walls)."""
    n = nx * ny * nz
    bodies = new_bodies(n)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    rng = np.random.default_rng(seed)          # jitter only: bulk generator for large N (seed recorded with the scene)
    j = rng.uniform(-jitter, jitter, size=(n, 3)).astype(np.float32)
    hx = F((nx - 1) * spacing * 0.5); hz = F((nz - 1) * spacing * 0.5)
    bodies["position"][:, 0] = ix.ravel().astype(np.float32) * F(spacing) - hx + j[:, 0]
    bodies["position"][:, 1] = iy.ravel().astype(np.float32) * F(spacing) + F(base_y) + j[:, 1]
    bodies["position"][:, 2] = iz.ravel().astype(np.float32) * F(spacing) - hz + j[:, 2]
    set_sphere_bodies(bodies, slice(None), radius, 1.0)
    ns = 5 if walls else 1
    colliders = np.zeros(n + ns, dtype=abi.collider_dtype)
    colliders["shape"][:n, 3] = radius
    colliders["restitution"][:n] = 0.1; colliders["friction"][:n] = 0.5; colliders["density"][:n] = 1.0
    colliders["type"][:n] = abi.SPHERE; colliders["objectType"][:n] = abi.OBJ_RIGID_BODY
    colliders["objectIndex"][:n] = np.arange(n, dtype=np.uint32); colliders["parentTransform"][:n] = abi.NO_TRANSFORM
    transforms = np.zeros(1, dtype=abi.transform_dtype)
    transforms["rotation"][0, 3] = 1.0
    ex = float(hx) + radius + 0.6; ez = float(hz) + radius + 0.6; h = float(ny * spacing) + 4.0
    static_box(colliders, n, (0, -4.0, 0), (ex + 8.0, 4.0, ez + 8.0), 0)
    if walls:
        static_box(colliders, n + 1, (-(ex + 2.0), h * 0.5, 0), (2.0, h * 0.5 + 2.0, ez + 4.0), 0)
        static_box(colliders, n + 2, ((ex + 2.0), h * 0.5, 0), (2.0, h * 0.5 + 2.0, ez + 4.0), 0)
        static_box(colliders, n + 3, (0, h * 0.5, -(ez + 2.0)), (ex + 4.0, h * 0.5 + 2.0, 2.0), 0)
        static_box(colliders, n + 4, (0, h * 0.5, (ez + 2.0)), (ex + 4.0, h * 0.5 + 2.0, 2.0), 0)
    return {"bodies": bodies, "colliders": colliders, "transforms": transforms, "name": "sphere_box_drop_%dx%dx%d" % (nx, ny, nz)}


def sphere_pile_settled(nx=16, ny=16, nz=16, spacing=0.96, jitter=0.01, seed=7, radius=0.5):
    """Dense overlapping lattice (every sphere touches its neighbours): exercises sphere-sphere contacts + colouring."""
    return sphere_box_drop(nx, ny, nz, spacing=spacing, jitter=jitter, seed=seed, radius=radius, base_y=0.48, walls=True)


def sphere_pile_slab(nx_total, ny, nz, ix0, ix1, spacing=0.96, jitter=0.01, seed=7, radius=0.5, base_y=0.48):
    """x-layers [ix0, ix1) of the (nx_total x ny x nz) settled sphere pile + the 5 static boxes of the WHOLE pile (replicated on
    every rank). Jitter is drawn per x-layer (seed, layer) so the scene is identical for every world size."""
    nx = ix1 - ix0
    n = nx * ny * nz
    bodies = new_bodies(n)
    hx = F((nx_total - 1) * spacing * 0.5); hz = F((nz - 1) * spacing * 0.5)
    iy, iz = np.meshgrid(np.arange(ny), np.arange(nz), indexing="ij")
    per = ny * nz
    for k, ix in enumerate(range(ix0, ix1)):
        j = np.random.default_rng(seed * 1000003 + ix).uniform(-jitter, jitter, size=(per, 3)).astype(np.float32)
        sl = slice(k * per, (k + 1) * per)
        bodies["position"][sl, 0] = F(ix) * F(spacing) - hx + j[:, 0]
        bodies["position"][sl, 1] = iy.ravel().astype(np.float32) * F(spacing) + F(base_y) + j[:, 1]
        bodies["position"][sl, 2] = iz.ravel().astype(np.float32) * F(spacing) - hz + j[:, 2]
    set_sphere_bodies(bodies, slice(None), radius, 1.0)
    colliders = np.zeros(n + 5, dtype=abi.collider_dtype)
    colliders["shape"][:n, 3] = radius
    colliders["restitution"][:n] = 0.1; colliders["friction"][:n] = 0.5; colliders["density"][:n] = 1.0
    colliders["type"][:n] = abi.SPHERE; colliders["objectType"][:n] = abi.OBJ_RIGID_BODY
    colliders["objectIndex"][:n] = np.arange(n, dtype=np.uint32); colliders["parentTransform"][:n] = abi.NO_TRANSFORM
    transforms = np.zeros(1, dtype=abi.transform_dtype); transforms["rotation"][0, 3] = 1.0
    ex = float(hx) + radius + 0.6; ez = float(hz) + radius + 0.6; h = float(ny * spacing) + 4.0
    static_box(colliders, n, (0, -4.0, 0), (ex + 8.0, 4.0, ez + 8.0), 0)
    static_box(colliders, n + 1, (-(ex + 2.0), h * 0.5, 0), (2.0, h * 0.5 + 2.0, ez + 4.0), 0)
    static_box(colliders, n + 2, ((ex + 2.0), h * 0.5, 0), (2.0, h * 0.5 + 2.0, ez + 4.0), 0)
    static_box(colliders, n + 3, (0, h * 0.5, -(ez + 2.0)), (ex + 4.0, h * 0.5 + 2.0, 2.0), 0)
    static_box(colliders, n + 4, (0, h * 0.5, (ez + 2.0)), (ex + 4.0, h * 0.5 + 2.0, 2.0), 0)
    x_lo = float(F(ix0) * F(spacing) - hx) - 0.5 * spacing; x_hi = float(F(ix1 - 1) * F(spacing) - hx) + 0.5 * spacing
    return {"bodies": bodies, "colliders": colliders, "transforms": transforms, "slab": (x_lo, x_hi), "name": "sphere_pile_slab_%d_%d_of_%d" % (ix0, ix1, nx_total)}


def event_zoo(per_type=6, seed=99):
    """Event-layer test scene: dynamic bodies with each of the six collider types fall (and are blown sideways) through a stack of
    six trigger volumes and six localized force fields, one per collider type, so that every ordered type pair of the boolean
    overlap tests (collision_narrow.cpp:1593-1689) occurs with a rigid body on one side. No static geometry: nothing collides
    except bodies with each other. Returns the scene plus `force_fields` (one world-space force per field) and `global_force`."""
    rng = np.random.default_rng(seed)
    # one shared hull geometry: an octahedron of radius 0.45
    hull = np.array([[0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0], [0, -0.45, 0], [0, 0, 0.45], [0, 0, -0.45]], dtype=np.float32)
    tris = np.array([0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5], dtype=np.uint32)
    types = [abi.SPHERE, abi.CAPSULE, abi.CYLINDER, abi.AABB, abi.OBB, abi.HULL]
    n = per_type * len(types)
    bodies = new_bodies(n)
    set_sphere_bodies(bodies, np.arange(n), F(0.4), F(1.0))
    volumes = 12
    colliders = np.zeros(n + volumes, dtype=abi.collider_dtype)
    colliders["restitution"] = 0.1; colliders["friction"] = 0.5; colliders["density"] = 1.0
    colliders["parentTransform"] = abi.NO_TRANSFORM

    def set_shape(i, t, size):
        s = colliders["shape"][i]; s[:] = 0
        if t == abi.SPHERE:
            s[3] = size
        elif t in (abi.CAPSULE, abi.CYLINDER):
            s[0:3] = (0, -0.5 * size, 0); s[3:6] = (0, 0.5 * size, 0); s[6] = 0.6 * size
        elif t == abi.AABB:
            s[0:3] = -size; s[3:6] = size
        elif t == abi.OBB:
            s[0:4] = (0, 0, 0, 1); s[7:10] = size
        else:
            s[0:4] = (0, 0, 0, 1); s[7:8].view(np.uint32)[0] = 0
        colliders["type"][i] = t

    for i in range(n):
        t = types[i // per_type]
        set_shape(i, t, 0.4)
        colliders["objectType"][i] = abi.OBJ_RIGID_BODY; colliders["objectIndex"][i] = i
        k = i % per_type
        bodies["position"][i] = (F(1.7) * (i // per_type) + rng.uniform(-0.2, 0.2), F(9.0) + F(1.1) * k, rng.uniform(-0.6, 0.6))
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(0, np.pi)
        if t != abi.AABB:            # an unrotated body keeps its AABB collider (physics.cpp:722-734)
            bodies["rotation"][i] = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]]).astype(np.float32)
        bodies["linearVelocity"][i] = (0, -4.0, 0)
        bodies["angularDamping"][i] = 0.0 if t == abi.AABB else 0.4
    transforms = np.zeros(volumes, dtype=abi.transform_dtype); transforms["rotation"][:, 3] = 1.0
    for v in range(volumes):
        i = n + v; t = types[v % 6]
        set_shape(i, t, 0.9)
        is_trigger = v < 6
        colliders["objectType"][i] = abi.OBJ_TRIGGER if is_trigger else abi.OBJ_FORCE_FIELD
        colliders["objectIndex"][i] = v % 6
        colliders["parentTransform"][i] = v
        # a long flat volume under all the columns would only exercise a few pairs: give every volume its own height and let it span x
        transforms["position"][v] = (F(1.7) * 2.5, F(6.5) - F(1.0) * (v % 6) - (0.0 if is_trigger else 0.5), 0)
        if t in (abi.SPHERE,):
            colliders["shape"][i][3] = 5.2
        elif t in (abi.CAPSULE, abi.CYLINDER):
            colliders["shape"][i][0:3] = (-4.6, 0, 0); colliders["shape"][i][3:6] = (4.6, 0, 0); colliders["shape"][i][6] = 0.8
        elif t == abi.AABB:
            colliders["shape"][i][0:3] = (-5.0, -0.4, -1.2); colliders["shape"][i][3:6] = (5.0, 0.4, 1.2)
        elif t == abi.OBB:
            colliders["shape"][i][7:10] = (5.0, 0.4, 1.2)
            q = np.array([0, 0, np.sin(0.05), np.cos(0.05)], dtype=np.float32); transforms["rotation"][v] = q
        else:
            transforms["position"][v][0] = F(1.7) * (v % 6)      # the small hull sits under one column only
    force_fields = np.zeros((6, 3), dtype=np.float32)
    for k in range(6):
        force_fields[k] = (3.0 + k, 0.5 * k, -2.0 + k)
    return {"bodies": bodies, "colliders": colliders, "transforms": transforms, "joints": {}, "name": "event_zoo",
            "hull_vertices": hull, "hull_offsets": np.array([0, len(hull)], dtype=np.uint32), "hull_triangles": [tris],
            "force_fields": force_fields, "global_force": np.array([0.0, 0.0, 0.75], dtype=np.float32)}


def contact_zoo(per_pair=16, seed=4242, spacing=4.0):
    """Narrow-phase test scene: for each of the 21 ordered collider type pairs (collision_narrow.cpp:374-1584) `per_pair` two-body
    instances that interpenetrate -- random relative poses plus the special configurations the routines branch on (parallel capsule /
    cylinder axes, cylinder cap-to-cap, segment shapes lying flat on a box face, axis-aligned boxes on every side, sphere centre inside
    a box, coincident sphere centres). Instances sit on a lattice `spacing` apart so that only the intended pairs touch. A body keeps
    an AABB collider only while its rotation is exactly identity (physics.cpp:722-734), so AABB bodies are never rotated here.
    No static geometry. Returns the usual scene dict; body 2k and 2k+1 form instance k."""
    rng = np.random.default_rng(seed)
    types = [abi.SPHERE, abi.CAPSULE, abi.CYLINDER, abi.AABB, abi.OBB, abi.HULL]
    pairs = [(a, b) for a in types for b in types if a <= b]
    n_inst = len(pairs) * per_pair
    n = 2 * n_inst
    bodies = new_bodies(n)
    set_sphere_bodies(bodies, np.arange(n), F(0.4), F(1.0))
    colliders = np.zeros(n, dtype=abi.collider_dtype)
    colliders["restitution"] = 0.1; colliders["friction"] = 0.5; colliders["density"] = 1.0
    colliders["parentTransform"] = abi.NO_TRANSFORM
    colliders["objectType"] = abi.OBJ_RIGID_BODY; colliders["objectIndex"] = np.arange(n, dtype=np.uint32)
    # two hull geometries: octahedron (6 vertices) and an irregular 8-vertex frustum
    octa = np.array([[0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0], [0, -0.45, 0], [0, 0, 0.45], [0, 0, -0.45]], dtype=np.float32)
    octa_t = np.array([0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5], dtype=np.uint32)
    fr = np.array([[-0.5, -0.4, -0.5], [0.5, -0.4, -0.5], [0.5, -0.4, 0.5], [-0.5, -0.4, 0.5], [-0.3, 0.4, -0.3], [0.3, 0.4, -0.3], [0.3, 0.4, 0.3], [-0.3, 0.4, 0.3]], dtype=np.float32)
    fr_t = np.array([0, 1, 2, 0, 2, 3, 4, 6, 5, 4, 7, 6, 0, 4, 5, 0, 5, 1, 1, 5, 6, 1, 6, 2, 2, 6, 7, 2, 7, 3, 3, 7, 4, 3, 4, 0], dtype=np.uint32)

    def rand_quat():
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(0, np.pi)
        return np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]]).astype(np.float32)

    def set_shape(i, t, k):
        s = colliders["shape"][i]; s[:] = 0
        if t == abi.SPHERE:
            s[0:3] = rng.uniform(-0.05, 0.05, 3); s[3] = rng.uniform(0.3, 0.5)
        elif t in (abi.CAPSULE, abi.CYLINDER):
            h = rng.uniform(0.25, 0.5); s[0:3] = (0, -h, 0); s[3:6] = (0, h, 0); s[6] = rng.uniform(0.15, 0.3)
        elif t == abi.AABB:
            r = rng.uniform(0.25, 0.5, 3); s[0:3] = -r; s[3:6] = r
        elif t == abi.OBB:
            s[0:4] = rand_quat() if (k & 1) else (0, 0, 0, 1); s[4:7] = rng.uniform(-0.05, 0.05, 3); s[7:10] = rng.uniform(0.25, 0.5, 3)
        else:
            s[0:4] = rand_quat() if (k & 1) else (0, 0, 0, 1); s[4:7] = rng.uniform(-0.05, 0.05, 3); s[7:8].view(np.uint32)[0] = k & 1
        colliders["type"][i] = t

    side = int(np.ceil(np.sqrt(n_inst)))
    inst = 0
    for (ta, tb) in pairs:
        for k in range(per_pair):
            ia, ib = 2 * inst, 2 * inst + 1
            base = np.array([(inst % side) * spacing, 5.0 + 0.37 * (inst % 3), (inst // side) * spacing], dtype=np.float32)
            set_shape(ia, ta, k); set_shape(ib, tb, k + 1)
            qa = np.array([0, 0, 0, 1], dtype=np.float32) if ta == abi.AABB else rand_quat()
            qb = np.array([0, 0, 0, 1], dtype=np.float32) if tb == abi.AABB else rand_quat()
            off = rng.normal(size=3); off = off / np.linalg.norm(off) * rng.uniform(0.05, 0.65)
            special = k % 8
            seg = (abi.CAPSULE, abi.CYLINDER)
            if ta in seg and tb in seg and special in (1, 2, 3):
                qb = qa.copy()                                  # parallel axes (|cos| > 0.99 branch)
                R = _quat_to_mat(qa)
                if special == 1: off = R @ np.array([0.3, 0.2, 0.0])            # side by side, overlapping intervals -> 2 contacts
                elif special == 2: off = R @ np.array([0.05, 0.75, 0.0])        # end to end (disjoint intervals / cap to cap)
                else: off = R @ np.array([0.25, 0.55, 0.1])                     # partial interval overlap
            elif ta in seg and tb in (abi.AABB, abi.OBB) and special in (1, 2):
                # segment shape lying flat on the +y face of the box: EPA normal along a box axis, axis perpendicular to it -> 2-point clip
                if tb == abi.OBB:
                    colliders["shape"][ib][0:4] = (0, 0, 0, 1)
                Rb = _quat_to_mat(qb)
                qa = _quat_mul(qb, np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)], dtype=np.float32))    # capsule axis along the box's x
                ry = colliders["shape"][ib][4] if tb == abi.AABB else colliders["shape"][ib][8]
                off = -(Rb @ np.array([0.02 * special, ry + colliders["shape"][ia][6] * 0.8, 0.03]))
            elif ta == abi.AABB and tb == abi.AABB:
                axis = special % 3; sign = 1.0 if (special & 1) else -1.0
                off = rng.uniform(-0.1, 0.1, 3); off[axis] = sign * rng.uniform(0.3, 0.6)                   # min-overlap axis with both signs (quirk B13)
            elif ta == abi.SPHERE and tb == abi.SPHERE and special == 1:
                off = np.zeros(3); colliders["shape"][ia][0:3] = 0; colliders["shape"][ib][0:3] = 0       # coincident centres -> normal (0,1,0)
            elif ta == abi.SPHERE and tb in (abi.AABB, abi.OBB) and special == 1:
                off = rng.uniform(-0.05, 0.05, 3)                                                          # sphere centre inside the box
            elif ta == abi.OBB and tb == abi.OBB and special in (1, 2):
                colliders["shape"][ia][0:4] = (0, 0, 0, 1); colliders["shape"][ib][0:4] = (0, 0, 0, 1)
                qb = qa.copy() if special == 1 else _quat_mul(qa, np.array([0, np.sin(0.3), 0, np.cos(0.3)], dtype=np.float32))
                off = _quat_to_mat(qa) @ np.array([0.1, 0.55, 0.05])                                       # face-face stacks (4-point manifolds, reduction)
            bodies["position"][ia] = base + 0.5 * off.astype(np.float32); bodies["position"][ib] = base - 0.5 * off.astype(np.float32)
            bodies["rotation"][ia] = qa; bodies["rotation"][ib] = qb
            bodies["linearVelocity"][ia] = rng.uniform(-1, 1, 3); bodies["linearVelocity"][ib] = rng.uniform(-1, 1, 3)
            bodies["angularVelocity"][ia] = rng.uniform(-1, 1, 3) if ta != abi.AABB else 0; bodies["angularVelocity"][ib] = rng.uniform(-1, 1, 3) if tb != abi.AABB else 0
            inst += 1
    transforms = np.zeros(1, dtype=abi.transform_dtype); transforms["rotation"][0, 3] = 1.0
    verts = np.concatenate([octa, fr]); offs = np.array([0, len(octa), len(octa) + len(fr)], dtype=np.uint32)
    return {"bodies": bodies, "colliders": colliders, "transforms": transforms, "joints": {}, "name": "contact_zoo_%d" % per_pair,
            "hull_vertices": verts, "hull_offsets": offs, "hull_triangles": [octa_t, fr_t], "pair_types": pairs, "per_pair": per_pair}


def _quat_to_mat(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _quat_mul(a, b):
    ax, ay, az, aw = [float(v) for v in a]; bx, by, bz, bw = [float(v) for v in b]
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], dtype=np.float32)


def terrain_heights(chunks_per_dim=2, chunk_size=16.0, amplitude=6.0, seed=5):
    """uint16 heights of a smooth rolling terrain, chunks_per_dim^2 chunks of 129 x 129 vertices (the reference's terrain chunk layout,
    terrain/heightmap_collider.h:8-11); neighbouring chunks share their border row / column."""
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 6.28, size=4)
    n = 129
    out = np.zeros((chunks_per_dim * chunks_per_dim, n, n), dtype=np.uint16)
    for cz in range(chunks_per_dim):
        for cx in range(chunks_per_dim):
            gx = (cx * 128 + np.arange(n))[None, :] * (chunk_size / 128.0)
            gz = (cz * 128 + np.arange(n))[:, None] * (chunk_size / 128.0)
            h = 0.45 + 0.22 * np.sin(0.35 * gx + ph[0]) * np.cos(0.27 * gz + ph[1]) + 0.12 * np.sin(0.9 * gx + 0.6 * gz + ph[2]) + 0.05 * np.cos(1.9 * gz + ph[3])
            out[cz * chunks_per_dim + cx] = np.clip(h * 65535.0, 0, 65535).astype(np.uint16)
    return out


def terrain_drop(n_side=10, seed=5, chunks_per_dim=2, chunk_size=16.0, amplitude=6.0):
    """Heightmap scene: n_side^2 bodies -- spheres, capsules, unrotated boxes (AABB colliders), rotated boxes (AABB colliders
    that become OBBs, physics.cpp:722-734) and OBB colliders -- start just above / slightly inside a rolling terrain and fall onto it. No
    other static geometry. The scene carries "heightmap": the arguments of b200phys_upload_heightmap."""
    rng = np.random.default_rng(seed)
    heights = terrain_heights(chunks_per_dim, chunk_size, amplitude, seed)
    extent = chunks_per_dim * chunk_size
    min_corner = np.array([-0.5 * extent, -2.0, -0.5 * extent], dtype=np.float32)
    n = n_side * n_side
    bodies = new_bodies(n)
    set_sphere_bodies(bodies, np.arange(n), F(0.4), F(1.0))
    colliders = np.zeros(n, dtype=abi.collider_dtype)
    colliders["restitution"] = 0.1; colliders["friction"] = 0.5; colliders["density"] = 1.0
    colliders["parentTransform"] = abi.NO_TRANSFORM
    colliders["objectType"] = abi.OBJ_RIGID_BODY; colliders["objectIndex"] = np.arange(n)

    def height(x, z):      # bilinear, like heightmap_collider_chunk::getHeightAt
        u = (x - min_corner[0]) / chunk_size; v = (z - min_corner[2]) / chunk_size
        cx = min(int(u), chunks_per_dim - 1); cz = min(int(v), chunks_per_dim - 1)
        fx = (u - cx) * 128.0; fz = (v - cz) * 128.0
        ix = min(int(fx), 127); iz = min(int(fz), 127); rx = fx - ix; rz = fz - iz
        hh = heights[cz * chunks_per_dim + cx].astype(np.float64) * (amplitude / 65535.0)
        a, b, c, d = hh[iz, ix], hh[iz + 1, ix], hh[iz, ix + 1], hh[iz + 1, ix + 1]
        return (a + rx * (c - a)) * (1 - rz) + (b + rx * (d - b)) * rz + min_corner[1]

    span = 0.7 * extent
    for i in range(n):
        kind = i % 5
        s = colliders["shape"][i]; s[:] = 0
        if kind == 0:
            colliders["type"][i] = abi.SPHERE; s[3] = 0.4
        elif kind == 1:
            colliders["type"][i] = abi.CAPSULE; s[0:3] = (0, -0.3, 0); s[3:6] = (0, 0.3, 0); s[6] = 0.25
        elif kind in (2, 3):
            colliders["type"][i] = abi.AABB; s[0:3] = (-0.35, -0.3, -0.4); s[3:6] = (0.35, 0.3, 0.4)
        else:
            colliders["type"][i] = abi.OBB; s[0:4] = (0, 0, 0, 1); s[4:7] = (0.05, 0.0, -0.05); s[7:10] = (0.3, 0.35, 0.25)
        x = -0.5 * span + span * ((i % n_side) + 0.5) / n_side + rng.uniform(-0.2, 0.2)
        z = -0.5 * span + span * ((i // n_side) + 0.5) / n_side + rng.uniform(-0.2, 0.2)
        bodies["position"][i] = (x, height(x, z) + rng.uniform(0.15, 0.7), z)
        if kind not in (0, 2):
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax); ang = rng.uniform(0, np.pi)
            bodies["rotation"][i] = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]]).astype(np.float32)
        bodies["linearVelocity"][i] = (rng.uniform(-0.5, 0.5), -2.0, rng.uniform(-0.5, 0.5))
        if kind == 2:
            bodies["angularDamping"][i] = 0.0
    transforms = np.zeros(1, dtype=abi.transform_dtype); transforms["rotation"][:, 3] = 1.0
    return {"bodies": bodies, "colliders": colliders, "transforms": transforms, "joints": {}, "name": "terrain_drop",
            "heightmap": dict(chunks_per_dim=chunks_per_dim, chunk_size=np.float32(chunk_size), min_corner=min_corner, amplitude_scale=np.float32(amplitude),
                              restitution=np.float32(0.1), friction=np.float32(0.9), heights=heights)}
