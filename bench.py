#!/usr/bin/env python
"""bench.py -- physics steps/sec of the B200-native rigid-body stepper (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W                    # our arm: the CUDA path through the C ABI (libb200phys.so)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's OWN code (oracle/_ref/libref_fast.so: its physics
                                                                   # sources compiled with g++, AVX2 path on) on the host cores

Headline workload (N = 1): BASELINE configs[2], 1 048 576 mixed primitives (sphere / capsule / OBB / convex hull round robin, random
orientations, 128 x 64 x 128 lattice settled into a pile inside 5 static boxes), 10 PGS iterations, dt 1/120. With N GPUs the pile is
(128 N) x 64 x 128, slab-decomposed along x, ghost-body halo between slab neighbours: weak scaling, 1 048 576 bodies per GPU, i.e. the
8.4 M-body scene of configs[4] at N = 8. A step = one physicsStepInternal (world colliders -> broad phase -> narrow phase -> PGS
iterations -> integration) over all bodies. `value` = bodies x steps/s with state resident in HBM; `e2e` = the same through the public
C-ABI calls with per-step pinned-host uploads of every pose / velocity / accumulator and the download of every pose / velocity.
The other single-GPU configurations (configs[0], [1], [3]) are measured in the same run and reported under "configs"; "like_for_like"
is configs[0] (1 024 spheres), the one scene the reference's 16-bit indices allow at full size, run by both arms.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line (NCCL prints its version banner to stdout)

import numpy as np

METRIC = "physics steps/sec at N rigid bodies (value = bodies x steps/s, whole job)"
ITERATIONS = 10
DT = np.float32(1.0 / 120.0)
MIXED_DIMS = (128, 64, 128)      # per GPU: 1 048 576 bodies (BASELINE configs[2])
MIXED_SPACING = 0.8
MIXED_SEED = 1578123
CPU_SAMPLE_DIMS = (32, 16, 32)   # 16 384 bodies of the same lattice: the largest the reference's uint16 collider indices allow (<= 32 767)
# algorithmic bytes per unit
B_GRID, B_PAIR_OUT = 136.0, 8.0          # broad phase: per collider (bin, sort, ranges, query re-read) + per emitted pair
B_PAIR_NARROW, B_CONTACT_OUT = 64.0, 40.0  # narrow phase: 2 x (16 shape + 16 tags) gathered per pair; contact 32 + body pair 8 written
B_KINIT = 316.0                          # row set-up: contact 40 + 2 x 76 body state read, 124-byte row written
B_KITER, B_JITER = 232.0, 376.0          # per contact / joint and PGS iteration
B_BODY = 424.0                           # force + velocity integration, world colliders


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the hot kernels from the committed `ncu --set full` summaries (profiles/r02_ncu_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------ scenes
def mixed_slab(rank, world, dims=MIXED_DIMS):
    """x-layers of rank `rank` of the (dims[0] * world) x dims[1] x dims[2] mixed pile + the static boxes of the whole pile."""
    from d3d12renderer_b200 import host
    nx = dims[0]
    return host.build_mixed_slab(nx * world, dims[1], dims[2], nx * rank, nx * (rank + 1), seed=MIXED_SEED, spacing=MIXED_SPACING)


def make_stepper(sc, device=0, **kw):
    from d3d12renderer_b200 import abi
    nb, nc = len(sc["bodies"]), len(sc["colliders"])
    nj = max([len(j[0]) for j in sc.get("joints", {}).values()] + [1])
    st = abi.Stepper(max_bodies=nb, max_colliders=nc, max_transforms=max(4, len(sc["transforms"])), max_joints=nj, device=device, **kw)
    if "hull_vertices" in sc:
        st.set_hulls(sc["hull_vertices"], sc["hull_offsets"])
    st.upload_transforms(np.ascontiguousarray(sc["transforms"])); st.upload_colliders(np.ascontiguousarray(sc["colliders"])); st.upload_bodies(np.ascontiguousarray(sc["bodies"]))
    for k, (j, p) in sc.get("joints", {}).items():
        st.upload_joints(k, j, p)
    return st


def kernel_table(st, do_step, reps):
    """Mean device time of every kernel group over `reps` steps (CUDA events on the launching stream, recorded by the library)."""
    st.enable_timing(True)
    acc = {}; order = []; counts = dict(contacts=0.0, pairs=0.0, manifolds=0.0); stage = np.zeros(7)
    for _ in range(reps):
        do_step()
        s = st.stats(check=False)
        stage += np.array([s.msWorld, s.msBroad, s.msNarrow, s.msInit, s.msSolve, s.msIntegrate, s.msTotal])
        counts["contacts"] += s.numContacts; counts["pairs"] += s.numBroadphaseOverlaps; counts["manifolds"] += s.numCollisions
        for name, ms in st.timeline():
            if name not in acc:
                acc[name] = 0.0; order.append(name)
            acc[name] += ms
    st.enable_timing(False)
    return {k: acc[k] / reps for k in order}, {k: v / reps for k, v in counts.items()}, stage / reps


def rooflines(kernels, counts, n_bodies, n_colliders, n_joints, peak):
    """Algorithmic bytes (SURVEY 8d model) / measured time for the kernel groups that move the data; -> {group: {...}}."""
    P, K = counts["pairs"], counts["contacts"]
    model = {
        "broad_phase": (n_colliders * B_GRID + P * B_PAIR_OUT, ["bp_check", "bp_build", "bp_tiles", "bp_emit", "bp_filter"]),
        "narrow_phase": (P * B_PAIR_NARROW + K * B_CONTACT_OUT, ["np_classify_sort", "np_closed_form_sat", "np_gjk", "np_epa", "np_epa_redo", "np_compact"]),
        "solver_rows": (K * B_KINIT, ["solver_rows"]),
        "solve": (K * ITERATIONS * B_KITER + n_joints * ITERATIONS * B_JITER, ["solve"]),
        "bodies": (n_bodies * B_BODY, ["world_colliders", "integrate_forces", "integrate_velocities"]),
    }
    out = {}
    for name, (nbytes, groups) in model.items():
        ms = sum(kernels.get(g, 0.0) for g in groups)
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        out[name] = {"algorithmic_bytes": nbytes, "ms": ms, "achieved": gbs, "frac": gbs / peak}
    return out


def time_steps(do_step, barrier, stream, warmup, steps):
    import torch
    for _ in range(warmup):
        do_step()
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        do_step()
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1)


def side_config(name, sc, peak, steps, warmup=5, iterations=ITERATIONS):
    """One of the other single-GPU BASELINE configurations, measured like the headline (device-resident)."""
    import torch
    from d3d12renderer_b200 import abi
    st = make_stepper(sc)
    stream = torch.cuda.current_stream()
    st.set_stream(stream.cuda_stream)
    sync = torch.cuda.synchronize
    ms = time_steps(lambda: st.step(DT, iterations, abi.STEP_FAST), sync, stream, warmup, steps)
    kernels, counts, stage = kernel_table(st, lambda: st.step(DT, iterations, abi.STEP_FAST), 6)
    ms_exact = time_steps(lambda: st.step(DT, iterations, abi.STEP_EXACT_ORDER), sync, stream, 3, max(4, steps // 4))
    s = st.stats(check=False)
    nj = sum(len(j[0]) for j in sc.get("joints", {}).values())
    rf = rooflines(kernels, counts, len(sc["bodies"]), len(sc["colliders"]), nj, peak)
    out = {"workload": name, "bodies": len(sc["bodies"]), "colliders": len(sc["colliders"]), "joints": nj, "contacts": counts["contacts"], "pairs": counts["pairs"],
           "ms_per_step": ms / steps, "steps_per_s": steps / (ms * 1e-3), "body_steps_per_s": len(sc["bodies"]) * steps / (ms * 1e-3),
           "ms_per_step_exact_order": ms_exact / max(4, steps // 4), "kernel_ms": kernels, "launches_per_step": int(s.kernelLaunches),
           "roofline": {k: {"frac": v["frac"], "achieved_gbs": v["achieved"], "ms": v["ms"]} for k, v in rf.items() if k in ("broad_phase", "solve")},
           "overflow": [int(s.overflowPairs), int(s.overflowManifolds), int(s.overflowGrid), int(s.epaOverflows)], "broadphase_reuse": st.broadphase_reuse()}
    st.close()
    return out


# --------------------------------------------------------------------------------------------------- the CPU arm
def reference_library():
    """(module, kind): the compiled reference (oracle/_ref, `reference`) when it is there, else the oracle restatement (`port`)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    if ref_lib.available(fast=True):
        flags = ""
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            pass
        if " avx2" in flags and " fma" in flags:
            return ref_lib, "reference"
    return None, "port"


def cpu_run(sc, steps, warmup, iterations=ITERATIONS):
    """The reference's own physicsStep (compiled from its sources, its AVX2 path on: physics.h:394-396 defaults) on `sc`, 1 thread like
    the reference's step; falls back to the oracle restatement (scalar path, -mavx2 -ffast-math build) when oracle/_ref is absent."""
    ref_lib, kind = reference_library()
    n = len(sc["bodies"])
    if kind == "reference":
        r = ref_lib.RefScene(sc, fast=True)
        for _ in range(warmup):
            r.step(DT, iterations, simd=True)
        t = 0.0
        for _ in range(steps):
            r.step(DT, iterations, simd=True); t += r.last_step_seconds()
        how = "the reference's own sources (src/physics, src/scene, EnTT) compiled with g++ -O2 -mavx2 -mfma -ffast-math (its /arch:AVX2 /fp:fast), physicsStep with simdBroadPhase = simdNarrowPhase = simdConstraintSolver = true (its defaults), 1 thread (the reference's step is single-threaded)"
    else:
        os.environ.setdefault("B200_ORACLE_VARIANT", "fast")
        import oracle_lib
        o = oracle_lib.Oracle(sc, canonical=False)
        for _ in range(warmup):
            o.step(DT, iterations)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.step(DT, iterations)
        t = time.perf_counter() - t0
        how = "oracle restatement of the reference's scalar path built -mavx2 -mfma -ffast-math (oracle/_ref not available here), 1 thread"
    return {"bodies": n, "steps": steps, "seconds": t, "ms_per_step": 1e3 * t / steps, "steps_per_s": steps / t, "body_steps_per_s": n * steps / t, "kind": kind, "how": how}


def cpu_sample_scene():
    from d3d12renderer_b200 import host
    return host.build_scene(host.CONFIG_MIXED_PILE, *CPU_SAMPLE_DIMS, seed=MIXED_SEED, spacing=MIXED_SPACING)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from d3d12renderer_b200 import scenes
    steps = max(1, min(args.steps, 60)); warmup = max(1, min(args.warmup, 2))
    sc = cpu_sample_scene()
    s = cpu_run(sc, steps, warmup)
    lfl = cpu_run(scenes.spheres_on_ground(32, mode="rest"), 200, 10)
    sample = ("%d-body mixed pile of the headline workload's lattice (%d x %d x %d of 128 x 64 x 128 per GPU; the reference's uint16 collider indices cap a scene at 32 767 colliders), "
              "%d steps in %.1f s; %s" % (s["bodies"], *CPU_SAMPLE_DIMS, steps, s["seconds"], s["how"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": s["body_steps_per_s"], "unit": "body-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d-body mixed-primitive pile (bounded SAMPLE of BASELINE configs[2]: same lattice, seed and spacing, %d x %d x %d), 10 PGS iterations, dt 1/120, measured as run (nothing extrapolated)" % (s["bodies"], *CPU_SAMPLE_DIMS),
                   "bodies": s["bodies"], "iterations": ITERATIONS, "steps_per_s": s["steps_per_s"], "sample_of": "1048576-body mixed pile per GPU"},
        "cpu_baseline": {"value": s["body_steps_per_s"], "unit": "body-steps/s", "cores": 1, "kind": s["kind"], "sample": sample},
        "like_for_like": {"workload": "BASELINE configs[0]: 1024 spheres resting on the ground box, 10 PGS iterations (full size on both arms)", "steps_per_s": lfl["steps_per_s"], "ms_per_step": lfl["ms_per_step"], "kind": lfl["kind"]},
        "e2e": {"value": s["body_steps_per_s"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timed-only", action="store_true", help="only the warm-up and the timed steps of the headline workload, then exit (ncu launch lists: nothing but the fast-order step is launched)")
    ap.add_argument("--no-side-configs", action="store_true", help="skip configs[0], [1], [3] and the strong-scaling point (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from d3d12renderer_b200 import abi, parallel, scenes, host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout to the one JSON line
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            t = torch.zeros(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    W = max(3, args.warmup); K = args.steps
    peak, peak_src = measured_hbm_peak()
    stream = torch.cuda.Stream()          # a real (non-default) stream: every kernel of the context and the timing events share it
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
        return x

    def sum_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())
        return x

    def fresh_nccl_id():
        """One ncclUniqueId per communicator (an id cannot be used twice): rank 0 creates it, torch.distributed hands it to the others."""
        box = [abi.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)          # the library owns its own NCCL communicator (b200phys_create_distributed)
        return box[0]

    def sharded(sc, ghost_layers=3.0):
        """Stepper for this rank's part of a slab-decomposed scene; with N > 1 the context is created by b200phys_create_distributed and
        b200phys_step_distributed exchanges the ghost halo with the slab neighbours (ncclSend / ncclRecv on the context's stream)."""
        n_bodies = len(sc["bodies"])
        per_layer = max(1, n_bodies // max(1, int(round((sc["slab"][1] - sc["slab"][0]) / MIXED_SPACING))))
        ghost_cap = 0 if world == 1 else int(ghost_layers * per_layer)
        planes = [None] * world
        if world > 1:
            dist.all_gather_object(planes, sc["slab"])
        else:
            planes = [sc["slab"]]
        bounds = np.array([-parallel.BIG] + [0.5 * (planes[r][1] + planes[r + 1][0]) for r in range(world - 1)] + [parallel.BIG])
        dc = None
        if world > 1:
            dc = abi.DistConfig(); dc.rank = rank; dc.nranks = world; dc.axis = 0; dc.slabLo = float(bounds[rank]); dc.slabHi = float(bounds[rank + 1]); dc.ghostWidth = 1.05
            dc.maxGhostBodies = ghost_cap; dc.maxGhostColliders = ghost_cap
        st = make_stepper(sc, device=local_rank, dist=(dc, fresh_nccl_id()) if dc is not None else None)
        st.set_stream(stream.cuda_stream)
        return st

    # ================================================================ headline: mixed pile, 1 048 576 bodies per GPU (weak scaling)
    sc = mixed_slab(rank, world)
    n_bodies = len(sc["bodies"]); n_col = len(sc["colliders"])
    st = sharded(sc)

    def do_step(flags=abi.STEP_FAST):
        if world > 1:
            st.step_distributed(DT, ITERATIONS, flags)      # pack (device) -> ncclSend / ncclRecv of fixed-size messages with the slab neighbours -> unpack (device) -> step; no host sync
        else:
            st.step(DT, ITERATIONS, flags)

    for _ in range(W):
        do_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_dev = max_over_ranks(time_steps(do_step, barrier, stream, 0, K))
    clock_info = clocks.stop() if rank == 0 else None
    stats = st.stats(check=False)
    reuse = st.broadphase_reuse()              # over warm-up + timed steps: how many of them ran the grid pipeline (the others filtered the candidate list)
    launches_per_step = stats.kernelLaunches
    overflow = [int(stats.overflowPairs), int(stats.overflowManifolds), int(stats.overflowGrid)]

    if args.timed_only:
        if rank == 0:
            print(json.dumps({"timed_only": True, "ms_per_step": ms_dev / K, "n_gpus": world, "steps": K, "warmup": W, "launches_per_step": int(launches_per_step), "overflow": overflow, "broadphase_reuse": reuse}))
        st.close()
        if world > 1:
            dist.destroy_process_group()
        return
    kernels, counts, stage = kernel_table(st, do_step, 8)
    rf = rooflines(kernels, counts, n_bodies, n_col, 0, peak)
    ms_exact = max_over_ranks(time_steps(lambda: do_step(abi.STEP_EXACT_ORDER), barrier, stream, 2, max(3, K // 5))) / max(3, K // 5)
    for _ in range(2):
        do_step()                                   # back on the fast order (colour memory) before the end-to-end loop

    # ---- end to end through the public API: H2D of every body state + accumulators, step, D2H of every pose / velocity ----------
    host_states = abi.PinnedArray(abi.body_state_dtype, n_bodies); hs = host_states.array
    host_forces = abi.PinnedArray(np.float32, n_bodies * 6); hf = host_forces.array.reshape(n_bodies, 6); hf[:] = 0
    st.halo_clear(); st.download_bodies(hs)
    for _ in range(2):
        st.upload_body_states(hs); st.upload_forces(hf); do_step(); st.halo_clear(); st.download_bodies(hs)
    barrier()
    K2 = max(5, K // 2)
    t0 = time.perf_counter()
    for _ in range(K2):
        st.upload_body_states(hs)                  # pinned host -> HBM: pose + velocities of every body, 52 B each
        st.upload_forces(hf)                       # pinned host -> HBM: force / torque accumulators, 24 B each (copy stream, overlaps collision detection)
        do_step()                                  # (+ ghost halo exchange with the slab neighbours when N > 1)
        st.halo_clear()
        st.download_bodies(hs)                     # HBM -> pinned host: pose + velocities of every OWNED body, 52 B each (synchronises)
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / K2
    checksum = float(np.sum(hs["position"][:, 1].astype(np.float64)))
    ghosts = st.dist_ghost_counts()[0] if world > 1 else 0
    st.close(); host_states.free(); host_forces.free()

    total_bodies = n_bodies * world
    value = total_bodies * K / (ms_dev * 1e-3)
    e2e_value = total_bodies / (ms_e2e * 1e-3)

    # ================================================================ strong scaling point (N > 1): ONE 1 048 576-body pile over N GPUs
    strong = None
    if world > 1 and not args.no_side_configs:
        nx = MIXED_DIMS[0] // world
        sc2 = host.build_mixed_slab(MIXED_DIMS[0], MIXED_DIMS[1], MIXED_DIMS[2], nx * rank, nx * (rank + 1), seed=MIXED_SEED, spacing=MIXED_SPACING)
        st2 = sharded(sc2)

        def step2():
            st2.step_distributed(DT, ITERATIONS, abi.STEP_FAST)
        ms2 = max_over_ranks(time_steps(step2, barrier, stream, W, K))
        st2.close()
        strong = {"workload": "ONE 1048576-body mixed pile (128 x 64 x 128) slab-decomposed over %d GPUs (fixed total: strong scaling)" % world, "bodies": MIXED_DIMS[0] * MIXED_DIMS[1] * MIXED_DIMS[2],
                  "ms_per_step": ms2 / K, "steps_per_s": K / (ms2 * 1e-3), "n_gpus": world, "note": "efficiency = (N=1 headline ms_per_step) / (N x this ms_per_step)"}

    # ================================================================ N = 1: the other single-GPU configurations + the CPU arm
    side = {}; lfl = None; cpu = None
    if rank == 0 and world == 1 and not args.no_side_configs:
        c1 = scenes.spheres_on_ground(32, mode="rest")
        side["configs[0]"] = side_config("1024 spheres resting on the ground box (BASELINE configs[0])", c1, peak, 200, warmup=20)
        side["configs[1]"] = side_config("262144-sphere settled pile in the open box (BASELINE configs[1])", scenes.sphere_pile_settled(64, 64, 64, seed=519431), peak, 40)
        side["configs[3]"] = side_config("16384 ragdolls: 229376 bodies, 278528 colliders, 114688 cone-twist + 98304 hinge joints (BASELINE configs[3])",
                                         host.build_scene(host.CONFIG_RAGDOLLS, 128, 1, 128, seed=20260924, mode=0, spacing=1.6), peak, 40)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = cpu_run(cpu_sample_scene(), 12, 1)
        cpu = {"value": s["body_steps_per_s"], "unit": "body-steps/s", "cores": 1, "kind": s["kind"], "ms_per_step": s["ms_per_step"],
               "sample": "%d-body mixed pile of the same lattice (%d x %d x %d; the reference's uint16 collider indices cap a scene at 32767 colliders), %d steps in %.1f s; %s"
                         % (s["bodies"], *CPU_SAMPLE_DIMS, s["steps"], s["seconds"], s["how"])}
        l = cpu_run(scenes.spheres_on_ground(32, mode="rest"), 200, 10)
        ours = side.get("configs[0]")
        lfl = {"workload": "BASELINE configs[0]: 1024 spheres resting on the ground box, 10 PGS iterations -- the one scene both arms run at full size",
               "reference_steps_per_s": l["steps_per_s"], "reference_ms_per_step": l["ms_per_step"], "reference_kind": l["kind"],
               "ours_steps_per_s": ours["steps_per_s"] if ours else None, "ours_ms_per_step": ours["ms_per_step"] if ours else None}

    if rank == 0:
        dom = max(kernels, key=lambda k: kernels[k])
        group_of = {"bp_check": "broad_phase", "bp_filter": "broad_phase", "bp_build": "broad_phase", "bp_tiles": "broad_phase", "bp_emit": "broad_phase", "np_classify_sort": "narrow_phase", "np_closed_form_sat": "narrow_phase", "np_gjk": "narrow_phase",
                    "np_epa": "narrow_phase", "np_epa_redo": "narrow_phase", "np_compact": "narrow_phase", "solver_rows": "solver_rows", "solver_order": "solver_rows", "solve": "solve"}
        # per-kernel algorithmic bytes for the dominant kernel group (same model, split by what that kernel touches)
        P, Kc = counts["pairs"], counts["contacts"]
        dom_bytes = {"solve": Kc * ITERATIONS * B_KITER, "bp_tiles": n_col * (24.0 + 128.0), "bp_emit": n_col * 128.0 + P * B_PAIR_OUT, "bp_build": n_col * (B_GRID - 24.0), "solver_rows": Kc * B_KINIT,
                     "np_epa": counts["manifolds"] * (152.0 + 128.0 + 80.0), "np_gjk": P * B_PAIR_NARROW, "np_closed_form_sat": P * B_PAIR_NARROW,
                     "np_compact": counts["manifolds"] * 80.0 + Kc * B_CONTACT_OUT + P * 16.0, "np_classify_sort": P * (8.0 + 32.0 + 1.0 + 4.0 + 16.0)}.get(dom, 0.0)
        dom_gbs = dom_bytes / (kernels[dom] * 1e-3) / 1e9 if kernels[dom] > 0 else 0.0
        traffic = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "1048576 mixed primitives per GPU (sphere / capsule / OBB / convex hull, BASELINE configs[2]; %d x 64 x 128 pile over %d GPU(s) = configs[4] layout), 10 PGS iterations, dt 1/120, fast (colour-sorted) solver order"
                                   % (MIXED_DIMS[0] * world, world),
                       "bodies": total_bodies, "bodies_per_gpu": n_bodies, "iterations": ITERATIONS, "contacts_per_gpu": counts["contacts"], "pairs_per_gpu": counts["pairs"], "manifolds_per_gpu": counts["manifolds"],
                       "steps_per_s": K / (ms_dev * 1e-3), "ms_per_step_exact_order": ms_exact,
                       "parallelism": ("slab decomposition along x, %d slabs, ghost-body halo: ncclSend / ncclRecv inside the C library (%d ghost bodies on rank 0)" % (world, ghosts)) if world > 1 else "single GPU",
                       "l2_policy": "per-step working set (rows, pair and manifold scratch, body planes: > 1 GB) >> 126 MB L2; no explicit flush",
                       "state_checksum_y": checksum, "overflow": overflow,
                       "broadphase_reuse": dict(reuse, note="warm-up + timed steps; 'rebuilds' of them ran the grid pipeline, the others filtered the candidate list of grown AABBs (same pair list either way)")},
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(n_bodies * (abi.body_state_dtype.itemsize + 24)),
                    "d2h_bytes_per_step": int(n_bodies * abi.body_state_dtype.itemsize)},
            "gpu_launches": int(launches_per_step * K),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak, "launch_ms": kernels[dom],
                         "algorithmic_bytes_per_launch": dom_bytes, "traffic": traffic.get(dom), "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/r02_ncu_traffic.json)" if traffic.get(dom) else None,
                         "peak_source": peak_src, "stage": group_of.get(dom),
                         "stages": {k: {"frac": v["frac"], "achieved": v["achieved"], "ms": v["ms"], "algorithmic_bytes": v["algorithmic_bytes"]} for k, v in rf.items()}},
            "kernel_ms": kernels,
            "stage_ms": {"world": stage[0], "broad": stage[1], "narrow": stage[2], "init": stage[3], "solve": stage[4], "integrate": stage[5], "total": stage[6]},
            "clocks": clock_info,
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if side:
            line["configs"] = side
        if lfl is not None:
            line["like_for_like"] = lfl
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
