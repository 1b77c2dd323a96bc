"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: slab partition, ghost regions, the neighbour exchange protocol
(counts then payloads) and scene sharding. The device-side pack / unpack kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d3d12renderer_b200 import abi, parallel, scenes


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        # every rank sends rank-specific records of rank-specific length to each slab neighbour
        send = {}
        for peer in ([rank - 1] if rank > 0 else []) + ([rank + 1] if rank < world - 1 else []):
            nb, nc = 3 + rank + 10 * peer, 2 + rank
            send[peer] = (torch.full((nb, parallel.BODY_RECORD_FLOATS), float(100 * rank + peer)), torch.full((nc, parallel.COLLIDER_RECORD_FLOATS), float(-(100 * rank + peer))))
        recv = parallel.exchange_with_neighbours(send, rank, world, dev)
        ok = True
        for peer, (b, c) in recv.items():
            ok &= b.shape == (3 + peer + 10 * rank, parallel.BODY_RECORD_FLOATS) and c.shape == (2 + peer, parallel.COLLIDER_RECORD_FLOATS)
            ok &= bool(torch.all(b == float(100 * peer + rank))) and bool(torch.all(c == float(-(100 * peer + rank))))
        # an empty message must work too
        send2 = {p: (torch.empty((0, parallel.BODY_RECORD_FLOATS)), torch.empty((0, parallel.COLLIDER_RECORD_FLOATS))) for p in send}
        recv2 = parallel.exchange_with_neighbours(send2, rank, world, dev)
        ok &= all(b.shape[0] == 0 and c.shape[0] == 0 for b, c in recv2.values())
        # the per-step protocol: one fixed-size message per neighbour and direction, header = counts, tail = padding
        words = 4 + 8 * parallel.BODY_RECORD_FLOATS + 8 * parallel.COLLIDER_RECORD_FLOATS
        out = {p: torch.full((words,), float(1000 * rank + p)) for p in send}
        inn = {p: torch.zeros(words) for p in send}
        parallel.exchange_fixed(out, inn)
        ok &= all(bool(torch.all(inn[p] == float(1000 * p + rank))) for p in inn)
        results[rank] = bool(ok) and set(recv.keys()) == set(send.keys())
    finally:
        dist.destroy_process_group()


def test_neighbour_exchange_world2_gloo():
    world = 2
    mgr = mp.Manager(); results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_neighbour_exchange_world3_gloo():
    world = 3   # the middle rank has two neighbours
    mgr = mp.Manager(); results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True, 2: True}


def test_slab_partition_and_ghost_regions():
    b = parallel.slab_bounds(4, -8.0, 8.0)
    assert len(b) == 5 and b[0] < -1e30 and b[-1] > 1e30 and np.allclose(b[1:4], (-4.0, 0.0, 4.0))
    x = np.array([-100.0, -4.0, -0.001, 0.0, 3.999, 4.0, 1e6])
    assert parallel.owner_of(x, b).tolist() == [0, 1, 1, 2, 2, 3, 3]
    assert parallel.ghost_regions(0, 4, b, 1.0) == [(1, -3.0 - 2.0, parallel.BIG)] or parallel.ghost_regions(0, 4, b, 1.0)[0][0] == 1
    r1 = parallel.ghost_regions(1, 4, b, 1.0)
    assert [p for p, _, _ in r1] == [0, 2] and r1[0][2] == -3.0 and r1[1][1] == -1.0
    assert parallel.ghost_regions(0, 1, parallel.slab_bounds(1, 0, 1), 1.0) == []


def test_shard_scene_covers_every_body_once_and_replicates_statics():
    sc = scenes.sphere_pile_settled(8, 4, 4)
    n = len(sc["bodies"])
    bounds = parallel.slab_bounds(2, float(sc["bodies"]["position"][:, 0].min()), float(sc["bodies"]["position"][:, 0].max()) + 1e-3)
    seen = np.zeros(n, dtype=int)
    for rank in range(2):
        sub, body_ids, col_ids = parallel.shard_scene(sc, rank, 2, bounds)
        seen[body_ids] += 1
        nb = len(sub["bodies"])
        assert np.array_equal(sub["bodies"], sc["bodies"][body_ids])
        dyn = sub["colliders"]["objectType"] == 0
        assert dyn.sum() == nb and np.all(sub["colliders"]["objectIndex"][dyn] < nb)
        assert (~dyn).sum() == 5          # floor + 4 walls on every rank
        # the remapped collider still points at the same global body
        assert np.array_equal(body_ids[sub["colliders"]["objectIndex"][dyn]], sc["colliders"]["objectIndex"][col_ids][dyn])
    assert np.all(seen == 1)


def test_articulated_groups_are_the_connected_components_of_the_joint_graph():
    pairs = np.zeros(4, dtype=abi.body_pair_dtype)
    pairs["rbA"] = [0, 1, 5, 7]; pairs["rbB"] = [1, 2, 6, 9]          # 0-1-2, 5-6, 7-(9 = the static dummy of an 9-body scene)
    g = parallel.articulated_groups(9, {3: (np.zeros(4), pairs)})
    assert g.tolist() == [0, 0, 0, 3, 4, 5, 5, 7, 8]


def test_shard_scene_keeps_jointed_mechanisms_whole_and_remaps_their_joints():
    """BASELINE config 4 (ragdolls) can be sharded: every ragdoll goes whole -- bodies, colliders, hinge and cone-twist joints -- to the slab that
    holds its mean position. With the ragdolls far enough apart not to touch each other the shards are independent islands on a shared static
    ground, so stepping each shard with the oracle must reproduce the full scene's body states BIT FOR BIT (the canonical contact and joint
    order of a sub-scene is the restriction of the full scene's order): that pins the body / collider / joint index remapping."""
    import oracle_lib
    from d3d12renderer_b200 import host
    sc = host.build_scene(host.CONFIG_RAGDOLLS, 4, 1, 3, seed=7, spacing=4.0)
    n = len(sc["bodies"]); assert n == 12 * 14
    xs = sc["bodies"]["position"][:, 0]
    full = oracle_lib.Oracle(sc, canonical=True)
    DT = np.float32(1.0 / 120.0)
    for _ in range(30):
        full.step(DT, 10)
    ref = full.body_states()
    for world in (2, 3):
        bounds = parallel.slab_bounds(world, float(xs.min()) - 0.5, float(xs.max()) + 0.5)
        seen = np.zeros(n, dtype=int); joints_seen = {k: 0 for k in sc["joints"]}
        for rank in range(world):
            sub, ids, cids = parallel.shard_scene(sc, rank, world, bounds)
            seen[ids] += 1
            assert len(ids) % 14 == 0 and np.array_equal(ids.reshape(-1, 14)[:, 0] % 14, np.zeros(len(ids) // 14, dtype=int))     # whole ragdolls, order kept
            nb = len(sub["bodies"])
            for k, (structs, pairs) in sub["joints"].items():
                joints_seen[k] += len(structs)
                assert np.all(pairs["rbA"] < nb) and np.all(pairs["rbB"] < nb)
                # mapped back to global ids every kept joint is one of the original joints, in the original order
                orig = sc["joints"][k][1]
                glob = np.stack([ids[pairs["rbA"]], ids[pairs["rbB"]]], 1)
                sel = np.nonzero(np.isin(orig["rbA"], ids))[0]
                assert np.array_equal(glob, np.stack([orig["rbA"][sel], orig["rbB"][sel]], 1)) and structs.tobytes() == sc["joints"][k][0][sel].tobytes()
            if nb == 0:
                continue
            orc = oracle_lib.Oracle(sub, canonical=True)
            for _ in range(30):
                orc.step(DT, 10)
            got = orc.body_states()
            assert got.tobytes() == ref[ids].tobytes(), "shard %d of %d does not reproduce the full scene" % (rank, world)
        assert np.all(seen == 1) and all(joints_seen[k] == len(sc["joints"][k][0]) for k in joints_seen)
