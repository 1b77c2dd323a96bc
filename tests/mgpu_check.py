"""Multi-GPU check (run with torchrun on >= 2 GPUs of one box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Slab-shards a sphere pile across the ranks with the ghost halo and compares with a single-GPU run of the whole scene:
 (1) the union over ranks of broad-phase pairs (mapped to global body ids) equals the single-GPU pair set,
 (2) after one step the sharded result differs from the single-GPU result by no more than Gauss-Seidel ORDER noise: 10 PGS
     iterations are far from converged on a pile, so two legal orders of the SAME scene (fast colouring vs canonical exact
     order) already differ at the 1e-2 level; the sharded run (different colouring + one-layer ghost approximation) must stay
     within 3x that noise, near the slab planes as well as far from them."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3d12renderer_b200 import abi, parallel, scenes  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    DT = np.float32(1 / 120.0)
    sc = scenes.sphere_pile_settled(8 * world, 8, 8, seed=11)
    n = len(sc["bodies"])
    xs = sc["bodies"]["position"][:, 0]
    bounds = parallel.slab_bounds(world, float(xs.min()) - 0.5, float(xs.max()) + 0.5)
    sub, body_ids, col_ids = parallel.shard_scene(sc, rank, world, bounds)
    cap_b = len(sub["bodies"]) + 2 * 512; cap_c = len(sub["colliders"]) + 2 * 512
    st = abi.Stepper(max_bodies=cap_b, max_colliders=cap_c, max_transforms=4, device=local)
    st.set_stream(stream.cuda_stream)
    st.upload_bodies(sub["bodies"]); st.upload_transforms(sub["transforms"]); st.upload_colliders(sub["colliders"])
    shard = parallel.SlabShard(st, rank, world, bounds, axis=0, ghost_width=1.05, max_ghost_bodies=512, max_ghost_colliders=512)
    shard.step(DT, 10, abi.STEP_FAST)
    stats = st.stats()
    states = st.download_bodies(np.empty(len(sub["bodies"]), dtype=abi.body_state_dtype))
    pairs = st.download_pairs()

    # local collider index -> global "collider key": dynamic colliders map to their global body id, statics to n + static index
    n_own_c = len(sub["colliders"])
    key = np.full(stats.numColliders, -1, dtype=np.int64)
    own = sub["colliders"]
    dyn = own["objectType"] == 0
    key[:n_own_c][dyn] = body_ids[own["objectIndex"][dyn]]
    key[:n_own_c][~dyn] = n + np.arange((~dyn).sum())
    # ghosts: fetch their owner-local indices from the packed records we received, map through the peers' global ids
    all_ids = [None] * world
    dist.all_gather_object(all_ids, body_ids)
    torch.cuda.synchronize()
    base = n_own_c; ghosts = [0, 0]
    for peer in sorted(shard.recv_msg.keys()):
        msg = shard.recv_msg[peer]
        hdr = msg[:2].view(torch.int32).cpu().numpy(); nb, ncol = int(hdr[0]), int(hdr[1])
        assert nb <= shard.cap_b and ncol <= shard.cap_c and nb == ncol    # one collider per ghost body in this scene, same order
        recs = msg[4:4 + shard.cap_b * 36].view(shard.cap_b, 36)
        owner_local = recs[:nb, 35].contiguous().view(torch.int32).cpu().numpy()
        key[base:base + ncol] = all_ids[peer][owner_local]
        base += shard.cap_c; ghosts[0] += nb; ghosts[1] += ncol          # the slots behind the counts are inert padding
    assert base == stats.numColliders
    ga = key[pairs["a"]]; gb = key[pairs["b"]]
    assert np.all(ga >= 0) and np.all(gb >= 0), "a padding slot took part in a pair"
    lo = np.minimum(ga, gb); hi = np.maximum(ga, gb)
    local_pairs = np.unique(np.stack([lo, hi], 1), axis=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, local_pairs)
    gathered_states = [None] * world
    dist.all_gather_object(gathered_states, (body_ids, states))

    if rank == 0:
        full = abi.Stepper(max_bodies=n, max_colliders=len(sc["colliders"]), max_transforms=4, device=local)
        full.upload_bodies(sc["bodies"]); full.upload_transforms(sc["transforms"]); full.upload_colliders(sc["colliders"])
        full.step(DT, 10, abi.STEP_FAST)
        ref_pairs = full.download_pairs(); ref_states = full.download_bodies()
        full2 = abi.Stepper(max_bodies=n, max_colliders=len(sc["colliders"]), max_transforms=4, device=local)
        full2.upload_bodies(sc["bodies"]); full2.upload_transforms(sc["transforms"]); full2.upload_colliders(sc["colliders"])
        full2.step(DT, 10, abi.STEP_EXACT_ORDER)
        alt_states = full2.download_bodies()
        c = sc["colliders"]
        ckey = np.where(c["objectType"] == 0, c["objectIndex"].astype(np.int64), n + np.cumsum(c["objectType"] != 0) - 1)
        ra = ckey[ref_pairs["a"]]; rb = ckey[ref_pairs["b"]]
        ref = np.unique(np.stack([np.minimum(ra, rb), np.maximum(ra, rb)], 1), axis=0)
        # static-static pairs exist on every rank; ghost-ghost pairs are duplicates of the owner's: the union must equal the reference set
        union = np.unique(np.concatenate(gathered, 0), axis=0)
        same = union.shape == ref.shape and np.array_equal(union, ref)
        print("pairs: single-GPU %d, union over %d ranks %d -> %s" % (len(ref), world, len(union), "IDENTICAL" if same else "DIFFERENT"))
        merged = np.empty(n, dtype=abi.body_state_dtype)
        for ids, s in gathered_states:
            merged[ids] = s
        planes = bounds[1:-1]
        dist_to_plane = np.min(np.abs(xs[:, None] - planes[None, :]), axis=1) if len(planes) else np.full(n, 1e9)
        def rel(x, f, mask):
            a = x[f][mask].astype(np.float64); b = ref_states[f][mask].astype(np.float64)
            return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0
        def rms(x, f, mask):
            a = x[f][mask].astype(np.float64); b = ref_states[f][mask].astype(np.float64)
            return float(np.sqrt(np.mean((a - b) ** 2))) if a.size else 0.0
        far = dist_to_plane > 3.0; near = ~far; everything = np.ones(n, dtype=bool)
        noise_max = rel(alt_states, "linearVelocity", everything); noise_rms = rms(alt_states, "linearVelocity", everything)
        print("order noise (single GPU, exact vs fast order): vel max rel %.3g rms %.3g" % (noise_max, noise_rms))
        print("sharded vs single GPU, far from planes (%d bodies): vel max rel %.3g rms %.3g pos max rel %.3g" % (far.sum(), rel(merged, "linearVelocity", far), rms(merged, "linearVelocity", far), rel(merged, "position", far)))
        print("sharded vs single GPU, within 3 m of a plane (%d bodies): vel max rel %.3g rms %.3g pos max rel %.3g" % (near.sum(), rel(merged, "linearVelocity", near), rms(merged, "linearVelocity", near), rel(merged, "position", near)))
        print("ghosts per rank (bodies, colliders):", tuple(ghosts), "owned:", len(sub["bodies"]))
        ok = same and rms(merged, "linearVelocity", everything) <= 3.0 * noise_rms and rel(merged, "linearVelocity", everything) <= 3.0 * noise_max
        print("MGPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
