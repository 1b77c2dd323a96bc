#!/usr/bin/env python
"""bench.py -- physics steps/sec of the B200-native rigid-body stepper on BASELINE.json's single-GPU headline configuration
(configs[1]: 262 144 spheres box-drop pile, 10 PGS iterations), one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W                   # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference's CPU algorithm (oracle restatement; the reference
                                                                  # itself is MSVC/Windows-only and does not compile here)
A step = one physicsStepInternal (world colliders -> broad phase -> narrow phase -> PGS iterations -> integration) over all bodies.
`value` is device-resident throughput (inputs already in HBM); `e2e` includes the per-step host->device upload of every
body record from pinned memory and the device->host download of every pose/velocity through the public C ABI calls.
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
# algorithmic bytes: per contact and solver iteration: row 120 + 2 x 28 read, 8 + 2 x 24 written
B_KITER = 232.0
# DRAM traffic of one solveKernel launch on the N=1 workload (778 k contacts, 10 iterations), from one `ncu --set full` capture:
# dram__bytes_read.sum 985.9 MB + dram__bytes_write.sum 11.6 MB (profiles/r01_final_solve_ncu_full_summary.txt). Below the
# algorithmic 1.80 GB because the 32-byte body records mostly hit in L2.
NCU_SOLVE_DRAM_BYTES = 985.903616e6 + 11.636224e6
B_GRID, B_PAIR_OUT = 136.0, 8.0
CPU_SAMPLE_DIMS = (32, 16, 32)   # 16 384 spheres: the largest pile the reference's uint16 collider indices allow (<= 32 767)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def build_scene(rank, world, nx=64):
    """configs[1]: 262 144 spheres per GPU (64 x 64 x 64 jittered lattice, spacing 0.96, r = 0.5: a settled pile with ~3 contacts
    per body) in an open box of 5 static AABBs. With N GPUs the pile is (64 N) x 64 x 64 and rank r owns x-layers [64 r, 64 (r+1)):
    ONE scene, slab-decomposed, coupled through the ghost-body halo (BASELINE configs[4] layout, weak scaling)."""
    from d3d12renderer_b200 import scenes
    return scenes.sphere_pile_slab(nx * world, 64, 64 if nx == 64 else 128, nx * rank, nx * (rank + 1), seed=519431)


def cpu_reference_sample(steps, warmup):
    """The reference's scalar CPU algorithm (oracle restatement) on a pile it can address (uint16 collider indices),
    same lattice and density as the GPU workload; 1 thread like the reference's step."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.setdefault("B200_ORACLE_VARIANT", "fast")   # timing build: -O2 -mavx2 -mfma -ffast-math like the reference's /arch:AVX2 /fp:fast
    import oracle_lib
    from d3d12renderer_b200 import scenes
    sc = scenes.sphere_pile_settled(*CPU_SAMPLE_DIMS, seed=519431)
    n = len(sc["bodies"])
    orc = oracle_lib.Oracle(sc, canonical=False)    # reference SAP order: no canonicalisation work on the clock
    for _ in range(warmup):
        orc.step(DT, ITERATIONS)
    stage = np.zeros(7)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.step(DT, ITERATIONS)
        stage += orc.stage_ms()
    t = time.perf_counter() - t0
    names = ["world_colliders", "broad_phase", "narrow_phase", "integrate_forces", "initialize_constraints", "solve_constraints", "integrate_velocities"]
    return {"bodies": n, "steps": steps, "seconds": t, "steps_per_s": steps / t, "body_steps_per_s": n * steps / t,
            "contacts": len(orc.contacts()[0]), "stage_ms": {k: float(v / steps) for k, v in zip(names, stage)}}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40)); warmup = max(1, min(args.warmup, 3))
    s = cpu_reference_sample(steps, warmup)
    total_bodies = 262144 * args.gpus
    line = {
        "impl": "reference", "metric": METRIC, "value": s["body_steps_per_s"], "unit": "body-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1000.0 * s["seconds"] / steps * (262144.0 / s["bodies"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "262144-sphere settled pile (BASELINE configs[1]), 10 PGS iterations, dt 1/120", "bodies": total_bodies, "iterations": ITERATIONS},
        "cpu_baseline": {"value": s["body_steps_per_s"], "unit": "body-steps/s", "cores": 1, "kind": "port",
                         "sample": "%d-sphere pile of the same lattice (the reference's uint16 indices cap colliders at 32767), %d steps, oracle restatement built -mavx2 -mfma -ffast-math (the reference's /arch:AVX2 /fp:fast), 1 thread like the reference's step; %d contacts; stage ms %s"
                                   % (s["bodies"], steps, s["contacts"], json.dumps(s["stage_ms"]))},
        "e2e": {"value": s["body_steps_per_s"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config5", action="store_true", help="1 048 576 spheres per GPU (128 x 64 x 128 slab per rank: BASELINE configs[4], 8.4 M bodies on 8 GPUs) instead of the default 262 144")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from d3d12renderer_b200 import abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(3, args.warmup); K = args.steps

    from d3d12renderer_b200 import parallel
    sc = build_scene(rank, world, 128 if args.config5 else 64)
    n_bodies = len(sc["bodies"]); n_col = len(sc["colliders"])
    ghost_cap = 0 if world == 1 else 3 * 64 * (128 if args.config5 else 64)      # ~2 layers of the slab face are ghosts, 1.5x head room
    st = abi.Stepper(max_bodies=n_bodies + 2 * ghost_cap, max_colliders=n_col + 2 * ghost_cap, max_transforms=4, device=local_rank)
    stream = torch.cuda.Stream()          # a real (non-default) stream: every kernel of the context and the timing events share it
    torch.cuda.set_stream(stream)
    st.set_stream(stream.cuda_stream)
    host_bodies = abi.PinnedArray(abi.body_dtype, n_bodies); host_bodies.array[:] = sc["bodies"]
    host_states = abi.PinnedArray(abi.body_state_dtype, n_bodies)
    st.upload_transforms(sc["transforms"]); st.upload_colliders(sc["colliders"]); st.upload_bodies(host_bodies.array)
    # slab planes between ranks; the ghost layer is one collider diameter (+5 %) wide
    x_lo, x_hi = sc["slab"]
    planes = [None] * world
    if world > 1:
        dist.all_gather_object(planes, (x_lo, x_hi))
    else:
        planes = [(x_lo, x_hi)]
    import numpy as _np
    bounds = _np.array([-parallel.BIG] + [0.5 * (planes[r][1] + planes[r + 1][0]) for r in range(world - 1)] + [parallel.BIG])
    shard = parallel.SlabShard(st, rank, world, bounds, axis=0, ghost_width=1.05, max_ghost_bodies=max(1, ghost_cap), max_ghost_colliders=max(1, ghost_cap),
                               device=torch.device("cuda", local_rank))

    def do_step():
        if world > 1:
            shard.refresh_ghosts()      # pack (device) -> one batched NCCL send/recv of fixed-size messages with the slab neighbours -> unpack (device); no host sync
        st.step(DT, ITERATIONS, abi.STEP_FAST)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
        return x

    # ---- device-resident throughput -----------------------------------------------------------------------------
    for _ in range(W):
        do_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        do_step()
    ev1.record(stream)
    barrier()
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1))
    clock_info = clocks.stop() if rank == 0 else None
    stats = st.stats()
    launches_per_step = stats.kernelLaunches

    # ---- per-kernel device times for the roofline (CUDA events on the launching stream, live) --------------------
    st.enable_timing(True)
    acc = np.zeros(7); contacts = 0; pairs = 0; cells = 0
    R = 10
    for _ in range(R):
        do_step()
        s = st.stats()
        acc += np.array([s.msWorld, s.msBroad, s.msNarrow, s.msInit, s.msSolve, s.msIntegrate, s.msTotal]); contacts += s.numContacts; pairs += s.numBroadphaseOverlaps
    st.enable_timing(False)
    acc /= R; contacts /= R; pairs /= R
    peak, peak_src = measured_hbm_peak()
    solve_bytes = contacts * ITERATIONS * B_KITER
    solve_gbs = solve_bytes / (acc[4] * 1e-3) / 1e9 if acc[4] > 0 else 0.0
    broad_bytes = n_col * B_GRID + pairs * B_PAIR_OUT
    broad_gbs = broad_bytes / (acc[1] * 1e-3) / 1e9 if acc[1] > 0 else 0.0

    # ---- end to end through the public API: H2D of every body record + step + D2H of every pose/velocity ----------
    hs = host_states.array
    host_forces = abi.PinnedArray(np.float32, n_bodies * 6); hf = host_forces.array.reshape(n_bodies, 6); hf[:] = 0
    st.download_bodies(hs)
    st.halo_clear()
    for _ in range(2):
        st.upload_body_states(hs); st.upload_forces(hf); do_step(); st.halo_clear(); st.download_bodies(hs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        st.upload_body_states(hs)                  # pinned host -> HBM: pose + velocities of every body, 52 B each
        st.upload_forces(hf)                       # pinned host -> HBM: force/torque accumulators, 24 B each
        do_step()                                  # (+ ghost halo exchange with the slab neighbours when N > 1)
        st.halo_clear()
        st.download_bodies(hs)                     # HBM -> pinned host: pose + velocities of every OWNED body, 52 B each (synchronises)
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    checksum = float(np.sum(hs["position"][:, 1].astype(np.float64)))

    total_bodies = n_bodies * world
    value = total_bodies * K / (ms_dev * 1e-3)
    e2e_value = total_bodies * K / (ms_e2e * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = cpu_reference_sample(20, 2)
        cpu = {"value": s["body_steps_per_s"], "unit": "body-steps/s", "cores": 1, "kind": "port",
               "sample": "%d-sphere pile of the same lattice (reference uint16 collider limit), %d steps in %.1f s, oracle restatement built -mavx2 -mfma -ffast-math (the reference's /arch:AVX2 /fp:fast), 1 thread like the reference's step" % (s["bodies"], s["steps"], s["seconds"])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("1048576-sphere settled pile per GPU (BASELINE configs[4])" if args.config5 else "262144-sphere settled pile per GPU (BASELINE configs[1])") + ", 10 PGS iterations, dt 1/120, fast (colour-sorted) solver order, dataflow sweep",
                       "bodies": total_bodies, "bodies_per_gpu": n_bodies, "iterations": ITERATIONS, "contacts_per_gpu": contacts, "pairs_per_gpu": pairs,
                       "steps_per_s": K / (ms_dev * 1e-3), "parallelism": ("slab decomposition along x, %d slabs, ghost-body halo over NCCL (%d ghost bodies on rank 0)" % (world, shard.ghost_counts()[0])) if world > 1 else "single GPU",
                       "l2_policy": "per-step working set (rows, pair and manifold scratch, body planes) > 126 MB L2; no explicit flush",
                       "state_checksum_y": checksum},
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": int(n_bodies * (abi.body_state_dtype.itemsize + 24)),
                    "d2h_bytes_per_step": int(n_bodies * abi.body_state_dtype.itemsize)},
            "gpu_launches": int(launches_per_step * K),
            "roofline": {"bound": "hbm", "kernel": "solveKernel (PGS iterations)", "achieved": solve_gbs, "peak": peak, "unit": "GB/s", "frac": solve_gbs / peak,
                         "traffic": NCU_SOLVE_DRAM_BYTES if world == 1 else None, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of solveKernel, ncu --set full on this workload (profiles/r01_final_solve_ncu_full_summary.txt)", "peak_source": peak_src, "algorithmic_bytes_per_launch": solve_bytes, "launch_ms": acc[4],
                         "broad_phase": {"achieved": broad_gbs, "frac": broad_gbs / peak, "algorithmic_bytes": broad_bytes, "ms": acc[1]}},
            "stage_ms": {"world": acc[0], "broad": acc[1], "narrow": acc[2], "init": acc[3], "solve": acc[4], "integrate": acc[5], "total": acc[6]},
            "clocks": clock_info,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    st.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
