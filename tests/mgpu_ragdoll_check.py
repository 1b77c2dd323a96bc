"""Multi-GPU check for jointed scenes (run with torchrun on >= 2 GPUs; NOT collected by pytest and NOT YET RUN ON HARDWARE: it was written
after the round's GPU budget was spent -- the host-side sharding it relies on is covered on the CPU by
tests/test_parallel_cpu.py::test_shard_scene_keeps_jointed_mechanisms_whole_and_remaps_their_joints).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_ragdoll_check.py

BASELINE config 4 in miniature: ragdolls (hinge + cone-twist chains) standing well apart on a ground box, sharded WHOLE across the ranks
(parallel.shard_scene), stepped with the library-owned halo transport (b200phys_step_distributed) in the exact solve order. The ragdolls
do not touch each other, so every rank's owned bodies are independent islands and must end BIT-IDENTICAL to a single-GPU run of the whole
scene; ragdolls near a slab plane reach the neighbour as ghosts (bodies and colliders, no joints) and must not disturb it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3d12renderer_b200 import abi, host, parallel  # noqa: E402


def upload(st, sc):
    st.upload_bodies(np.ascontiguousarray(sc["bodies"])); st.upload_transforms(np.ascontiguousarray(sc["transforms"])); st.upload_colliders(np.ascontiguousarray(sc["colliders"]))
    for kind, (joints, pairs) in sc.get("joints", {}).items():
        st.upload_joints(kind, joints, pairs)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    DT = np.float32(1 / 120.0); STEPS = 20
    sc = host.build_scene(host.CONFIG_RAGDOLLS, 4 * world, 1, 4, seed=7, spacing=4.0)
    n = len(sc["bodies"])
    xs = sc["bodies"]["position"][:, 0]
    bounds = parallel.slab_bounds(world, float(xs.min()) - 0.5, float(xs.max()) + 0.5)
    sub, body_ids, _ = parallel.shard_scene(sc, rank, world, bounds)
    nb, nc, cap = len(sub["bodies"]), len(sub["colliders"]), 1024
    nj = max([len(j[0]) for j in sub["joints"].values()] + [1])
    ids = [abi.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dc = abi.DistConfig(); dc.rank = rank; dc.nranks = world; dc.axis = 0; dc.slabLo = float(bounds[rank]); dc.slabHi = float(bounds[rank + 1]); dc.ghostWidth = 2.5
    dc.maxGhostBodies = cap; dc.maxGhostColliders = cap
    st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=max(1, len(sub["transforms"])), max_joints=nj, device=local, dist=(dc, ids[0]))
    upload(st, sub)
    for _ in range(STEPS):
        st.step_distributed(DT, 30, abi.STEP_EXACT_ORDER)
    gb, gc = st.dist_ghost_counts()
    mine = st.download_bodies(np.empty(nb, dtype=abi.body_state_dtype))
    stats = st.stats(check=False)
    gathered = [None] * world
    dist.all_gather_object(gathered, (body_ids, mine, gb, gc, int(stats.overflowGrid)))
    if rank == 0:
        full = abi.Stepper(max_bodies=n, max_colliders=len(sc["colliders"]), max_transforms=max(1, len(sc["transforms"])), max_joints=max(len(j[0]) for j in sc["joints"].values()), device=local)
        upload(full, sc)
        for _ in range(STEPS):
            full.step(DT, 30, abi.STEP_EXACT_ORDER)
        ref = full.download_bodies()
        ok = True; seen = np.zeros(n, dtype=int)
        for r, (ids_r, states, g1, g2, ovf) in enumerate(gathered):
            seen[ids_r] += 1
            same = states.tobytes() == ref[ids_r].tobytes()
            print("rank %d: %d bodies (%d ragdolls), ghosts %d bodies / %d colliders, overflow flags %d: %s" % (r, len(ids_r), len(ids_r) // 14, g1, g2, ovf, "BIT-IDENTICAL to the single-GPU run" if same else "DIFFERENT"))
            ok = ok and same and ovf == 0
        ok = ok and bool(np.all(seen == 1)) and sum(g[2] for g in gathered) > 0
        print("MGPU_RAGDOLL_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
