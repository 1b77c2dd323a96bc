"""Multi-GPU check of the broad phase's candidate reuse (run with torchrun on >= 2 GPUs; launched by tests/test_gpu_multi.py).
Two distributed contexts per rank (each with its own NCCL communicator) step the same slab of a pile that comes to rest: one keeps the
candidate list of grown AABBs between steps, the other runs the grid pipeline in every step (b200phys_config.reserved[2] = 1). The ghost
slots behind the owned colliders are rewritten by every halo exchange, so the containment check sees them like any other collider. Pair
lists (ghosts included) and owned body states must be BIT-IDENTICAL in every step, and the list must actually have been reused."""
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
    DT = np.float32(1 / 120.0)
    sc = scenes.sphere_pile_settled(8 * world, 8, 8, seed=11)
    xs = sc["bodies"]["position"][:, 0]
    bounds = parallel.slab_bounds(world, float(xs.min()) - 0.5, float(xs.max()) + 0.5)
    sub, body_ids, _ = parallel.shard_scene(sc, rank, world, bounds)
    nb, nc, cap = len(sub["bodies"]), len(sub["colliders"]), 512

    def make(reuse):
        ids = [abi.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dc = abi.DistConfig(); dc.rank = rank; dc.nranks = world; dc.axis = 0; dc.slabLo = float(bounds[rank]); dc.slabHi = float(bounds[rank + 1]); dc.ghostWidth = 1.05
        dc.maxGhostBodies = cap; dc.maxGhostColliders = cap
        st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=4, device=local, dist=(dc, ids[0]), bp_reuse=reuse)
        st.upload_bodies(sub["bodies"]); st.upload_transforms(sub["transforms"]); st.upload_colliders(sub["colliders"])
        return st

    a, b = make(True), make(False)
    steps, same = 120, True
    first_bad = -1
    for step in range(steps):
        a.step_distributed(DT, 10, abi.STEP_FAST); b.step_distributed(DT, 10, abi.STEP_FAST)
        if step % 4 == 3 or step < 8:
            pa, pb = a.download_pairs(), b.download_pairs()
            sa = a.download_bodies(np.empty(nb, dtype=abi.body_state_dtype)); sb = b.download_bodies(np.empty(nb, dtype=abi.body_state_dtype))
            ok = np.array_equal(pa, pb) and sa.tobytes() == sb.tobytes()
            if not ok and same:
                same = False; first_bad = step
    ra, rb = a.broadphase_reuse(), b.broadphase_reuse()
    gb, gc = a.dist_ghost_counts()
    out = [None] * world
    dist.all_gather_object(out, (bool(same), first_bad, ra, rb["steps"], gb, gc))
    if rank == 0:
        good = True
        for r, (s, bad, reuse, plain_steps, g1, g2) in enumerate(out):
            print("rank %d: reuse vs rebuild over %d distributed steps: %s%s; ghosts %d / %d; reuse context: %s" % (r, steps, "BIT-IDENTICAL" if s else "DIFFERENT", "" if s else " (first at step %d)" % bad, g1, g2, reuse))
            good = good and s and g1 > 0 and plain_steps == 0 and reuse["steps"] == steps and reuse["rebuilds"] < steps
        print("MGPU_REUSE_CHECK", "PASS" if good else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
