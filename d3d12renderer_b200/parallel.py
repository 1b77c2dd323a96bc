"""Multi-GPU plumbing: bodies sharded by spatial slab, one process (and one b200phys context) per GPU, ghost-body halo exchange
between slab neighbours with torch.distributed point-to-point operations (NCCL over NVLink / NVSwitch on the GPU box, gloo
in the CPU tests). The data path has exactly one exchange per step (two neighbours per rank); there is no global collective.

Ownership: rank r owns the bodies whose centre lies in [bounds[r], bounds[r+1]) along `axis` at scene-build time. Before every
step a rank packs the owned bodies whose collider AABBs reach within `ghost_width` of a neighbour's slab (device kernels,
k_halo.cu), sends them, and appends what it receives as ghosts. Ghost results are discarded after the step (owner wins).
"""
import numpy as np
import torch
import torch.distributed as dist

BODY_RECORD_FLOATS = 36       # b200phys_halo_record_bytes(0) / 4
COLLIDER_RECORD_FLOATS = 16   # b200phys_halo_record_bytes(1) / 4
BIG = 3.0e38


def slab_bounds(world, lo, hi):
    """world+1 plane positions splitting [lo, hi] into equal slabs; the outermost planes are open (-inf / +inf)."""
    b = np.linspace(lo, hi, world + 1).astype(np.float64)
    b[0] = -BIG; b[-1] = BIG
    return b


def owner_of(x, bounds):
    """Rank owning coordinate(s) x: the slab [bounds[r], bounds[r+1]) that contains it."""
    return np.clip(np.searchsorted(bounds, x, side="right") - 1, 0, len(bounds) - 2)


def ghost_regions(rank, world, bounds, ghost_width):
    """[(peer, lo, hi)]: owned bodies whose AABB overlaps [lo, hi) along the axis are ghosts for `peer`."""
    out = []
    if rank > 0:
        out.append((rank - 1, -BIG, float(bounds[rank]) + ghost_width))
    if rank < world - 1:
        out.append((rank + 1, float(bounds[rank + 1]) - ghost_width, BIG))
    return out


def exchange_with_neighbours(send, rank, world, device, group=None):
    """send: {peer: (bodies_tensor [nb, 36] float32, colliders_tensor [nc, 16] float32)} -> received {peer: (bodies, colliders)}.
    Two rounds of point-to-point operations: counts, then payloads sized by the counts (NCCL on CUDA tensors, gloo on CPU)."""
    peers = sorted(send.keys())
    if not peers:
        return {}
    cnt_out = {p: torch.tensor([send[p][0].shape[0], send[p][1].shape[0]], dtype=torch.int64, device=device) for p in peers}
    cnt_in = {p: torch.zeros(2, dtype=torch.int64, device=device) for p in peers}
    ops = []
    for p in peers:
        ops.append(dist.P2POp(dist.isend, cnt_out[p], p, group)); ops.append(dist.P2POp(dist.irecv, cnt_in[p], p, group))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    recv = {}
    ops = []
    for p in peers:
        nb, nc = int(cnt_in[p][0].item()), int(cnt_in[p][1].item())
        recv[p] = (torch.empty((nb, BODY_RECORD_FLOATS), dtype=torch.float32, device=device), torch.empty((nc, COLLIDER_RECORD_FLOATS), dtype=torch.float32, device=device))
        for t_out, t_in in ((send[p][0], recv[p][0]), (send[p][1], recv[p][1])):
            if t_out.numel():
                ops.append(dist.P2POp(dist.isend, t_out.contiguous(), p, group))
            if t_in.numel():
                ops.append(dist.P2POp(dist.irecv, t_in, p, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return recv


def exchange_fixed(send, recv, group=None):
    """One round of point-to-point operations on FIXED-size messages: send[peer] -> peer, recv[peer] <- peer (tensors of the same
    size on both sides). On CUDA tensors (NCCL) wait() only orders torch's current stream behind the transfer: no host sync."""
    ops = []
    for p in sorted(send.keys()):
        ops.append(dist.P2POp(dist.isend, send[p], p, group)); ops.append(dist.P2POp(dist.irecv, recv[p], p, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


class SlabShard:
    """One rank's shard: wraps a Stepper whose uploaded bodies are the OWNED ones and keeps its ghost layer fresh.

    Per step and neighbour one fixed-size message ([counts][cap body records][cap collider records], b200phys_halo_pack_message)
    goes each way; the receiver appends all slots, the unused ones as inert padding, so no count ever travels through the host:
    pack kernels -> one batched NCCL send/recv -> unpack kernel, all stream-ordered. The Stepper's stream must be torch's
    current stream (`torch.cuda.set_stream`) so that NCCL orders itself against the pack and unpack kernels."""

    def __init__(self, stepper, rank, world, bounds, axis=0, ghost_width=1.05, max_ghost_bodies=1 << 14, max_ghost_colliders=1 << 14, device=None):
        self.st = stepper; self.rank = rank; self.world = world; self.bounds = bounds; self.axis = axis; self.ghost_width = ghost_width
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.regions = ghost_regions(rank, world, bounds, ghost_width)
        self.cap_b = int(max_ghost_bodies); self.cap_c = int(max_ghost_colliders)
        words = stepper.halo_message_bytes(self.cap_b, self.cap_c) // 4
        self.send_msg = {p: torch.zeros(words, dtype=torch.float32, device=self.device) for p, _, _ in self.regions}
        self.recv_msg = {p: torch.zeros(words, dtype=torch.float32, device=self.device) for p, _, _ in self.regions}

    @staticmethod
    def measure(stepper, rank, world, bounds, axis=0, ghost_width=1.05):
        """(bodies, colliders) the largest ghost region of this rank would send right now (synchronous; for sizing the messages)."""
        nb = nc = 0
        for _, lo, hi in ghost_regions(rank, world, bounds, ghost_width):
            b, c = stepper.halo_pack(axis, lo, hi, 0, 0, 0x7FFFFFFF, 0x7FFFFFFF)
            nb = max(nb, b); nc = max(nc, c)
        return nb, nc

    def refresh_ghosts(self):
        """halo_clear -> pack per neighbour -> neighbour exchange -> unpack; asynchronous, call before every step."""
        st = self.st
        st.halo_clear()
        if self.world == 1:
            return
        for peer, lo, hi in self.regions:
            st.halo_pack_message(self.axis, lo, hi, self.send_msg[peer].data_ptr(), self.cap_b, self.cap_c)
        exchange_fixed(self.send_msg, self.recv_msg)
        for peer in sorted(self.recv_msg.keys()):
            st.halo_unpack_message(self.recv_msg[peer].data_ptr(), self.cap_b, self.cap_c)

    def migrate(self, ids=None):
        """Hands bodies that left this rank's slab (position beyond a slab plane) over to the neighbouring rank, colliders included;
        every rank must call it at the same step (every few steps is enough: until then the old owner keeps simulating the
        body, so keep ghost_width >= collider size + the distance a body travels between two migrations). Synchronous.
        ids: optional per-owned-body int64 array (e.g. global ids) that is carried along; returns the new array."""
        st = self.st
        st.halo_clear()
        if self.world == 1:
            return ids
        if not hasattr(self, "mig_send"):
            words = st.halo_message_bytes(self.cap_b, self.cap_c) // 4
            self.mig_send = [torch.zeros(words, dtype=torch.float32, device=self.device) for _ in range(2)]       # down, up
            self.mig_recv = {p: torch.zeros(words, dtype=torch.float32, device=self.device) for p, _, _ in self.regions}
        lo, hi = float(self.bounds[self.rank]), float(self.bounds[self.rank + 1])
        dest, counts = st.migrate_pack(self.axis, lo, hi, self.mig_send[0].data_ptr(), self.mig_send[1].data_ptr(), self.cap_b, self.cap_c)
        send = {}
        if self.rank > 0:
            send[self.rank - 1] = self.mig_send[0]
        if self.rank < self.world - 1:
            send[self.rank + 1] = self.mig_send[1]
        exchange_fixed(send, self.mig_recv)
        received = {}
        for peer in sorted(self.mig_recv.keys()):
            received[peer] = st.migrate_unpack(self.mig_recv[peer].data_ptr(), self.cap_b, self.cap_c)
        self.last_migration = (counts, received)
        if ids is None:
            return None
        ids = np.asarray(ids)
        out_ids = {}
        if self.rank > 0:
            out_ids[self.rank - 1] = torch.from_numpy(np.ascontiguousarray(ids[dest == 1]).astype(np.int64)).to(self.device)
        if self.rank < self.world - 1:
            out_ids[self.rank + 1] = torch.from_numpy(np.ascontiguousarray(ids[dest == 2]).astype(np.int64)).to(self.device)
        in_ids = {p: torch.empty(received[p][0], dtype=torch.int64, device=self.device) for p in received}   # sizes known from the messages
        ops = []
        for p in sorted(out_ids.keys()):
            if out_ids[p].numel():
                ops.append(dist.P2POp(dist.isend, out_ids[p], p))
            if in_ids[p].numel():
                ops.append(dist.P2POp(dist.irecv, in_ids[p], p))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        return np.concatenate([ids[dest == 0]] + [in_ids[p].cpu().numpy() for p in sorted(in_ids.keys())])

    def ghost_counts(self):
        """(bodies, colliders) received in the last exchange, summed over neighbours (reads the message headers: synchronous)."""
        tb = tc = 0
        for p in self.recv_msg:
            h = self.recv_msg[p][:2].view(torch.int32).cpu()
            tb += int(h[0]); tc += int(h[1])
        return tb, tc

    def step(self, dt, iterations, flags=0):
        self.refresh_ghosts()
        self.st.step(dt, iterations, flags)


def articulated_groups(n_bodies, joints):
    """Connected components of the joint graph: group[b] = smallest body index of b's mechanism (bodies without joints are their own group).
    `joints`: {kind: (structs, pairs)} as in the scene dicts; pairs carry rbA / rbB. A joint to a body index >= n_bodies (the static dummy)
    connects nothing."""
    parent = np.arange(n_bodies, dtype=np.int64)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]; x = parent[x]
        return x
    for _, (_, pairs) in sorted(joints.items()):
        for a, b in zip(pairs["rbA"].tolist(), pairs["rbB"].tolist()):
            if a >= n_bodies or b >= n_bodies:
                continue
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    return np.array([find(i) for i in range(n_bodies)], dtype=np.int64)


def shard_scene(scene, rank, world, bounds, axis=0):
    """Splits a single-GPU scene dict into this rank's shard: owned bodies with their colliders, static colliders replicated on every rank.
    A body without joints is owned by the slab that holds its centre. Jointed mechanisms (ragdolls, vehicles: the connected components of the
    joint graph) are sharded WHOLE: every body of a mechanism goes, with all its joints, to the slab that holds the mechanism's mean position,
    so no joint ever spans two ranks (a limb that sticks out into the neighbour's slab is still owned here and reaches the neighbour as a ghost,
    where it takes part in contacts only). Joint body indices are remapped to the shard; a joint's partner >= len(bodies) (the static dummy)
    becomes the shard's own dummy index. Returns (shard_scene, global_body_ids, global_collider_ids)."""
    bodies, colliders = scene["bodies"], scene["colliders"]
    n = len(bodies)
    joints = {k: v for k, v in (scene.get("joints") or {}).items() if len(v[0])}
    x = bodies["position"][:, axis].astype(np.float64)
    if joints:
        group = articulated_groups(n, joints)
        sums = np.bincount(group, weights=x, minlength=n); counts = np.bincount(group, minlength=n)
        x = (sums / np.maximum(counts, 1))[group]                 # every body votes with its mechanism's mean position
    owner = owner_of(x, bounds)
    mine = np.nonzero(owner == rank)[0]
    remap = np.full(n, -1, dtype=np.int64); remap[mine] = np.arange(len(mine))
    is_body = colliders["objectType"] == 0
    keep = np.nonzero(~is_body | (remap[np.minimum(colliders["objectIndex"], n - 1)] >= 0))[0]
    sub_c = colliders[keep].copy()
    bsel = sub_c["objectType"] == 0
    sub_c["objectIndex"][bsel] = remap[sub_c["objectIndex"][bsel]].astype(np.uint32)
    sub_j = {}
    for k, (structs, pairs) in joints.items():
        a = pairs["rbA"].astype(np.int64); b = pairs["rbB"].astype(np.int64)
        anchor = np.where(a < n, a, b)                                # the dynamic end decides (both ends share a mechanism otherwise)
        sel = np.nonzero((anchor < n) & (remap[np.minimum(anchor, n - 1)] >= 0))[0]
        p = pairs[sel].copy()
        for f, src in (("rbA", a), ("rbB", b)):
            v = src[sel]
            loc = np.where(v < n, remap[np.minimum(v, n - 1)], len(mine))
            assert np.all(loc >= 0), "a joint spans two mechanisms' owners"      # cannot happen: connected components are sharded whole
            p[f] = loc.astype(p[f].dtype)
        sub_j[k] = (np.ascontiguousarray(structs[sel]), np.ascontiguousarray(p))
    out = {"bodies": np.ascontiguousarray(bodies[mine]), "colliders": np.ascontiguousarray(sub_c), "transforms": scene["transforms"], "joints": sub_j}
    for k in ("hull_vertices", "hull_offsets", "hull_triangles"):
        if k in scene:
            out[k] = scene[k]
    return out, mine, keep
