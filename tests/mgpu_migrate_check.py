"""Multi-GPU migration check (run with torchrun on >= 2 GPUs of one box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_migrate_check.py

A sphere pile that drifts along +x is slab-sharded; bodies cross the slab planes and are handed over every few steps.
Checks: (1) after every migration each body is owned by exactly one rank and lies in that rank's slab (up to the drift of the
steps since the last migration); (2) at the end the sharded state is rebuilt as ONE single-GPU scene (gathered poses and
velocities) and both take one more step: the union of the ranks' broad-phase pairs equals the single-GPU pair set exactly;
(3) nothing blew up (finite, energy of the same order as a single-GPU run of the whole trajectory)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3d12renderer_b200 import abi, parallel, scenes  # noqa: E402

STEPS, EVERY, DRIFT_V = 40, 5, 4.0


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    DT = np.float32(1 / 120.0)
    sc = scenes.sphere_pile_settled(8 * world, 6, 8, seed=23)
    sc["bodies"]["linearVelocity"][:, 0] = DRIFT_V
    n = len(sc["bodies"])
    xs = sc["bodies"]["position"][:, 0]
    bounds = parallel.slab_bounds(world, float(xs.min()) - 0.5, float(xs.max()) + 0.5)
    sub, body_ids, col_ids = parallel.shard_scene(sc, rank, world, bounds)
    cap = 1024
    st = abi.Stepper(max_bodies=n + 2 * cap, max_colliders=len(sc["colliders"]) + 2 * cap, max_transforms=4, device=local)
    st.set_stream(stream.cuda_stream)
    st.upload_bodies(sub["bodies"]); st.upload_transforms(sub["transforms"]); st.upload_colliders(sub["colliders"])
    gw = 1.05 + DRIFT_V * 1.5 * EVERY / 120.0          # collider size + what a body can travel between two migrations
    shard = parallel.SlabShard(st, rank, world, bounds, axis=0, ghost_width=gw, max_ghost_bodies=cap, max_ghost_colliders=cap)
    ids = body_ids.astype(np.int64)
    # host mirror of the collider list (global key per collider): kept colliders keep their order, immigrants are appended
    c = sub["colliders"]
    ckeys = np.where(c["objectType"] == 0, body_ids[np.minimum(c["objectIndex"], len(body_ids) - 1)].astype(np.int64), n + np.cumsum(c["objectType"] != 0) - 1)
    moved_total = 0; ok = True
    for step in range(STEPS):
        if step and step % EVERY == 0:
            old_ids = ids
            ids = shard.migrate(ids)
            counts, received = shard.last_migration
            gone = set(np.setdiff1d(old_ids, ids).tolist())
            ckeys = np.concatenate([ckeys[[k not in gone for k in ckeys.tolist()]], ids[len(old_ids) - len(gone):]])
            moved_total += counts[0] + counts[2]
            every = [None] * world
            dist.all_gather_object(every, ids)
            allids = np.concatenate(every)
            ok &= len(allids) == n and len(np.unique(allids)) == n
            pos = st.download_bodies(np.empty(len(ids), dtype=abi.body_state_dtype))["position"][:, 0]
            ok &= bool(np.all(pos >= bounds[rank] - 1e-6) and np.all(pos < bounds[rank + 1] + 1e-6)) if len(pos) else True
        shard.step(DT, 10, abi.STEP_FAST)
    # hand over once more so that ownership is current, then compare one step with a single-GPU scene rebuilt from the gathered state
    old_ids = ids
    ids = shard.migrate(ids)
    gone = set(np.setdiff1d(old_ids, ids).tolist())
    ckeys = np.concatenate([ckeys[[k not in gone for k in ckeys.tolist()]], ids[len(old_ids) - len(gone):]])
    states = st.download_bodies(np.empty(len(ids), dtype=abi.body_state_dtype))
    gathered_states = [None] * world
    dist.all_gather_object(gathered_states, (ids, states))
    shard.step(DT, 10, abi.STEP_FAST)
    stats = st.stats()
    pairs = st.download_pairs()
    torch.cuda.synchronize()
    all_ids = [None] * world
    dist.all_gather_object(all_ids, ids)
    key = np.full(stats.numColliders, -1, dtype=np.int64)
    key[:len(ckeys)] = ckeys
    base = len(ckeys)
    for peer in sorted(shard.recv_msg.keys()):
        msg = shard.recv_msg[peer]
        hdr = msg[:2].view(torch.int32).cpu().numpy(); nb, ncol = int(hdr[0]), int(hdr[1])
        recs = msg[4:4 + shard.cap_b * 36].view(shard.cap_b, 36)
        owner_local = recs[:nb, 35].contiguous().view(torch.int32).cpu().numpy()
        key[base:base + ncol] = all_ids[peer][owner_local]
        base += shard.cap_c
    ga = key[pairs["a"]]; gb = key[pairs["b"]]
    ok &= bool(np.all(ga >= 0) and np.all(gb >= 0))
    local_pairs = np.unique(np.stack([np.minimum(ga, gb), np.maximum(ga, gb)], 1), axis=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, local_pairs)
    oks = [None] * world
    dist.all_gather_object(oks, (bool(ok), int(moved_total)))
    if rank == 0:
        bodies = sc["bodies"].copy()
        for gids, s in gathered_states:
            for f in ("rotation", "position", "linearVelocity", "angularVelocity"):
                bodies[f][gids] = s[f]
        full = abi.Stepper(max_bodies=n, max_colliders=len(sc["colliders"]), max_transforms=4, device=local)
        full.upload_bodies(bodies); full.upload_transforms(sc["transforms"]); full.upload_colliders(sc["colliders"])
        full.step(DT, 10, abi.STEP_FAST)
        ref_pairs = full.download_pairs()
        cc = sc["colliders"]
        ck = np.where(cc["objectType"] == 0, cc["objectIndex"].astype(np.int64), n + np.cumsum(cc["objectType"] != 0) - 1)
        ra = ck[ref_pairs["a"]]; rb = ck[ref_pairs["b"]]
        ref = np.unique(np.stack([np.minimum(ra, rb), np.maximum(ra, rb)], 1), axis=0)
        union = np.unique(np.concatenate(gathered, 0), axis=0)
        same = union.shape == ref.shape and np.array_equal(union, ref)
        moved = sum(m for _, m in oks)
        finite = all(np.all(np.isfinite(s["position"])) and np.all(np.isfinite(s["linearVelocity"])) for _, s in gathered_states)
        print("migrated bodies over %d steps: %d ; ownership invariants on every rank: %s ; finite: %s" % (STEPS, moved, all(o for o, _ in oks), finite))
        print("pairs after the last hand-over: single-GPU %d, union over %d ranks %d -> %s" % (len(ref), world, len(union), "IDENTICAL" if same else "DIFFERENT"))
        print("MGPU_MIGRATE_CHECK", "PASS" if (same and finite and moved > 0 and all(o for o, _ in oks)) else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
