"""Multi-GPU check of the library-owned transport (run with torchrun on >= 2 GPUs; launched by tests/test_gpu_multi.py):
b200phys_create_distributed / b200phys_step_distributed (NCCL inside libb200phys.so, include/b200phys.h) against the same slab
decomposition driven from Python (parallel.SlabShard: the identical pack / unpack kernels, torch.distributed as transport).
Both exchange the same fixed-size ghost messages, so after several steps the owned body states must be BIT-IDENTICAL."""
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
    xs = sc["bodies"]["position"][:, 0]
    bounds = parallel.slab_bounds(world, float(xs.min()) - 0.5, float(xs.max()) + 0.5)
    sub, body_ids, _ = parallel.shard_scene(sc, rank, world, bounds)
    nb, nc, cap = len(sub["bodies"]), len(sub["colliders"]), 512

    # (a) Python-driven halo
    st = abi.Stepper(max_bodies=nb + 2 * cap, max_colliders=nc + 2 * cap, max_transforms=4, device=local)
    st.set_stream(stream.cuda_stream)
    st.upload_bodies(sub["bodies"]); st.upload_transforms(sub["transforms"]); st.upload_colliders(sub["colliders"])
    shard = parallel.SlabShard(st, rank, world, bounds, axis=0, ghost_width=1.05, max_ghost_bodies=cap, max_ghost_colliders=cap)
    for _ in range(5):
        shard.step(DT, 10, abi.STEP_FAST)
    st.halo_clear()
    a = st.download_bodies(np.empty(nb, dtype=abi.body_state_dtype))

    # (b) the library's own NCCL transport
    ids = [abi.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dc = abi.DistConfig(); dc.rank = rank; dc.nranks = world; dc.axis = 0; dc.slabLo = float(bounds[rank]); dc.slabHi = float(bounds[rank + 1]); dc.ghostWidth = 1.05
    dc.maxGhostBodies = cap; dc.maxGhostColliders = cap
    st2 = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=4, device=local, dist=(dc, ids[0]))
    st2.upload_bodies(sub["bodies"]); st2.upload_transforms(sub["transforms"]); st2.upload_colliders(sub["colliders"])
    for _ in range(5):
        st2.step_distributed(DT, 10, abi.STEP_FAST)
    gb, gc = st2.dist_ghost_counts()
    st2.halo_clear()
    b = st2.download_bodies(np.empty(nb, dtype=abi.body_state_dtype))
    same = all(np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)) for f in a.dtype.names)
    flags = [None] * world
    dist.all_gather_object(flags, (bool(same), gb, gc, shard.ghost_counts()))
    if rank == 0:
        for r, (s, g1, g2, pg) in enumerate(flags):
            print("rank %d: library transport vs Python transport after 5 steps: %s; ghosts (bodies, colliders) %d, %d (Python path %s)" % (r, "BIT-IDENTICAL" if s else "DIFFERENT", g1, g2, pg))
        print("MGPU_CABI_CHECK", "PASS" if all(s for s, _, _, _ in flags) and all(g1 > 0 for _, g1, _, _ in flags) else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
