#!/usr/bin/env python
"""Benchmark of the point -> occupancy hot path (BASELINE.json metric: tracklets/s and voxel-steps/s
of the occupancy ray-cast, %HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

A *step* is one pass of the whole annotate path (crop -> box frame -> voxelise -> range-image
visibility) over one batch of synthetic tracklets.  The N=1 workload is BASELINE.json configs[1]:
64 vehicle tracklets x 40 frames at 0.2 m voxels, one shared segment of 5-LiDAR range images.
For N>1 (torchrun, one rank per GPU) every rank annotates its own batch of that shape (tracklets are
independent: no collective on the data path, "weak" scaling); rank 0 prints ONE JSON line.

`value`     tracklets/s with the batch resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`       the same through the public API with HOST buffers: per step, H2D of all inputs from pinned
            memory + the kernels + D2H of labels/dims/status; steps alternate between two device buffer sets
            on two streams, so one step's upload overlaps the previous step's kernels (events around all K).
`roofline`  the visibility ("ray-cast") kernel: algorithmic bytes 4*U*B*L + 4*V per tracklet
             / its mean launch duration, measured live with CUDA events
            recorded around that kernel on the launching stream (occb200_profile_*).
`cpu_baseline` the CPU oracle port (oracle/occ_oracle.c, OpenMP over tracklets) on the same workload,
            on this host's cores.  `--impl reference` times that port as the reference arm: the
            reference's own implementation of this path is Python+torch ops that cannot be imported
            here (mmcv/mmdet asserts, argparse at import); the port restates it op for op and is
            pinned bit-exactly against it (oracle/validate_oracle.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": "c1: 1 vehicle tracklet x 20 frames, 0.2 m voxels, 5 LiDARs",
    "c2": "c2: 64 vehicle tracklets x 40 frames, 0.2 m voxels, 5 LiDARs (one shared segment)",
    "c3": "c3: 16 truck/bus tracklets x 40 frames, 0.1 m voxels, 5 LiDARs",
    "c5s": "c5 (scaled): 1024 vehicle tracklets x 40 frames over 16 segments, 0.2 m voxels, 5 LiDARs",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-threads", type=int, default=0, help="0 = all host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-f64", action="store_true", help="all-f64 visibility kernel")
    ap.add_argument("--flags", type=int, default=0, help="extra occb200_annotate_args_t.flags bits (A/B measurements)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes per launch of the ray-cast kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "vis_fast_ncu.json")
    try:
        m = json.load(open(p))["metrics"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return sum(float(m[k]["value"]) * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_stats(res, B, L):
    """Nominal work of a batch from its results: voxel-steps U*B*L and algorithmic bytes 4*U*B*L + 4*V."""
    U = sum(r["n_unknown"] for r in res if r["occ"] is not None)
    V = sum(int(r["occ"].size) for r in res if r["occ"] is not None)
    steps = U * B * L
    return dict(U=U, V=V, steps=steps, vis_bytes=4 * steps + 4 * V,
                executed=sum(r.get("n_steps", 0) for r in res if r["occ"] is not None))


# ------------------------------------------------------------------------------------------------
def cpu_port(batch, threads, min_seconds=8.0, max_reps=3):
    """Time the CPU oracle port on the whole batch with `threads` OpenMP threads (packing excluded)."""
    from oracle import oracle

    pk = oracle.PackedBatch(batch)
    times, res = [], None
    t_end = time.perf_counter() + min_seconds
    while len(times) < max_reps and (not times or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        res = oracle.annotate_batch(batch, threads=threads, packed=pk)
        times.append(time.perf_counter() - t0)
    return min(times), res, len(times)


def run_reference(args):
    """--impl reference: the CPU port on all host cores, same config/metric/unit; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from objectcentricocccompletion_b200 import synth

    threads = args.cpu_threads or os.cpu_count()
    batch = synth.config_batch(args.workload, seed=0)
    B, L = len(batch.tracklets[0]), len(batch.segments[0].inclinations)
    from oracle import oracle

    pk = oracle.PackedBatch(batch)
    for _ in range(min(args.warmup, 1)):
        oracle.annotate_batch(batch, threads=threads, packed=pk)
    steps = max(1, min(args.steps, 5))          # bounded: each step is the full batch (seconds of CPU work)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = oracle.annotate_batch(batch, threads=threads, packed=pk)
    dt = (time.perf_counter() - t0) / steps
    ws = workload_stats(res, B, L)
    T = len(batch.tracklets)
    val = T / dt
    line = {"impl": "reference", "metric": "tracklets_per_s", "value": val, "unit": "tracklets/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the same workload keys as the CUDA arm's config (one batch per step; the CPU arm runs it on rank 0 only)
            "config": {"workload": WORKLOADS[args.workload], "tracklets_per_step_per_gpu": T, "frames": B, "lidars": L,
                       "voxel_size": batch.voxel_size,
                       "ok_tracklets": sum(r["occ"] is not None for r in res)},
            "voxel_steps_per_s": ws["steps"] / dt,
            "cpu_baseline": {"value": val, "unit": "tracklets/s", "cores": threads, "kind": "port",
                             "sample": f"full {args.workload} batch ({T} tracklets) per step, {steps} steps, OpenMP over tracklets"},
            "e2e": {"value": val, "unit": "tracklets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import objectcentricocccompletion_b200 as occ
    from objectcentricocccompletion_b200 import _lib, occ_annotate, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: one batch per rank (different seed per rank) -------------------------------
    batch = synth.config_batch(args.workload, seed=rank)
    T = len(batch.tracklets)
    B, L = len(batch.tracklets[0]), len(batch.segments[0].inclinations)
    pk = occ_annotate.pack_tracklets(batch)
    host = occ_annotate.HostBuffers(pk, pin=True)
    d = occ_annotate.DeviceTracklets(pk, dev)
    flags = (occ_annotate.FLAG_FORCE_F64 if args.force_f64 else 0) | args.flags
    d.upload(host)
    d.run(flags)
    torch.cuda.synchronize()
    res = d.results()
    n_recheck, q_cap = d.queue_stats(flags)
    ws = workload_stats(res, B, L)
    n_ok = sum(r["occ"] is not None for r in res)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, n):
        evs = []
        for _ in range(n):
            flush.zero_()                                              # evict L2 between steps (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)                  # ms

    # pinned result buffers for the e2e leg
    out_host = {k: torch.empty_like(getattr(d, k), device="cpu").pin_memory() for k in ("labels", "dims", "status", "n_unknown")}

    use_graph = not args.no_graph
    graph_kernels = d.capture(flags) if use_graph else 0

    def step_resident():
        if use_graph:
            d.replay(flags)
        else:
            d.run(flags)

    def step_e2e():
        d.upload(host)
        d.run(flags)
        for k, h in out_host.items():
            h.copy_(getattr(d, k), non_blocking=True)

    h2d_bytes = host.nbytes()
    d2h_bytes = sum(h.numel() * h.element_size() for h in out_host.values())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = graph_kernels * args.steps if use_graph else _lib.launch_count() - l0
    barrier()

    # ---- roofline pass: the same steps with events around each kernel ---------------------------
    _lib.lib().occb200_profile_enable(1)
    timed(lambda: d.run(flags), args.steps)           # direct launches: the events sit between the kernels
    _lib.lib().occb200_profile_enable(0)
    nk = _lib.lib().occb200_profile_kinds()
    kms = np.zeros(nk, np.float64)
    kn = np.zeros(nk, np.int64)
    _lib.check(_lib.lib().occb200_profile_read(kms.ctypes.data, kn.ctypes.data), "occb200_profile_read")

    # ---- e2e: host buffers -> H2D -> kernels -> D2H every step, double-buffered on two streams so the upload
    # of step k+1 overlaps the kernels / download of step k (inputs per step 134 MB > L2: no flush needed) -------
    d2 = occ_annotate.DeviceTracklets(pk, dev)
    if use_graph:
        d2.upload(host)
        d2.capture(flags)
    out_host2 = {k: torch.empty_like(h).pin_memory() for k, h in out_host.items()}
    pipes = [(torch.cuda.Stream(dev), d, out_host), (torch.cuda.Stream(dev), d2, out_host2)]

    def e2e_run(n):
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        for st, _, _ in pipes:
            st.wait_event(start)
        for i in range(n):
            st, dd, oh = pipes[i % 2]
            with torch.cuda.stream(st):
                dd.upload(host)
                if use_graph:
                    dd.replay(flags)
                else:
                    dd.run(flags)
                for k, h in oh.items():
                    h.copy_(getattr(dd, k), non_blocking=True)
        ends = []
        for st, _, _ in pipes:
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            ends.append(e)
        torch.cuda.synchronize()
        return max(start.elapsed_time(e) for e in ends)

    e2e_run(min(args.warmup, 3) + 1)
    barrier()
    ms_e2e = e2e_run(args.steps)
    barrier()
    assert all(bool((out_host2[k] == out_host[k]).all()) for k in out_host), "pipelined e2e results differ"
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(T), float(ws["steps"]), float(ws["executed"])], dtype=torch.float64, device=dev)
    gather_ms = None
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        # the one collective of the job: final gather of per-rank results to rank 0 over NVLink (not per step)
        from objectcentricocccompletion_b200 import dist as occ_dist

        occ_dist.gather_results(res[:1], [rank], world, dst=0)      # opens the NCCL p2p channels (lazy, ~1 s)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        allres = occ_dist.gather_results(res, list(range(rank * T, (rank + 1) * T)), world * T, dst=0)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            assert sum(r is not None and r["occ"] is not None for r in allres) >= n_ok
    ms, ms_e2e = float(t[0]), float(t[1])
    T_all, steps_all, exec_all = (float(x) for x in tot)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    sec = ms / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    peak, peak_src = peaks()
    vis_ms = (kms[4] + kms[3]) / max(kn[4], 1)            # the ray-cast: k_brick_cull + k_visibility
    achieved = ws["vis_bytes"] / (vis_ms / 1e3) / 1e9 if vis_ms > 0 else 0.0
    step_ms_prof = kms.sum() / max(kn[4], 1)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = args.cpu_threads or os.cpu_count()
        dt, cres, reps = cpu_port(batch, threads)
        mism = sum(int((c["occ"] != g["occ"]).sum()) for c, g in zip(cres, res) if c["occ"] is not None)
        cpu = {"value": T / dt, "unit": "tracklets/s", "cores": threads, "kind": "port",
               "sample": f"full {args.workload} batch ({T} tracklets), best of {reps}, OpenMP over tracklets, "
                         f"C port of the reference path; labels vs GPU: {mism} mismatches",
               "voxel_steps_per_s": ws["steps"] / dt}

    line = {
        "metric": "tracklets_per_s", "value": T_all / sec, "unit": "tracklets/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "tracklets_per_step_per_gpu": T, "frames": B, "lidars": L,
                   "voxel_size": batch.voxel_size, "ok_tracklets": n_ok, "l2": "flushed (256 MiB write) between timed steps",
                   "visibility": "f64" if args.force_f64 else "default",
                   "launch": "cuda graph replay" if use_graph else "kernel by kernel"},
        "voxel_steps_per_s": steps_all / sec, "executed_steps_per_s": exec_all / sec,
        "voxel_steps_per_step": ws["steps"], "executed_steps_per_step": ws["executed"],
        "f64_rechecks_per_step": n_recheck, "unknown_voxels_per_step": ws["U"], "voxels_per_step": ws["V"],
        "e2e": {"value": T_all / sec_e2e, "unit": "tracklets/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": sec_e2e * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_visibility", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic() if args.workload == "c2" else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ws["vis_bytes"], "kernel_ms": vis_ms,
                     "kernel_share_of_step": vis_ms / step_ms_prof if step_ms_prof else None,
                     "kernels_ms": dict(zip(("k_crop_voxelize", "k_tracklet_setup+redo", "k_scan_chunks", "k_brick_cull",
                                             "k_visibility", "side:k_pyr_build+k_table_setup",
                                             "k_visibility_recheck", "k_pair_build", "k_labels"),
                                            (kms / np.maximum(kn[4], 1)).round(5).tolist()))},
        "cpu_baseline": cpu, "clocks": clocks,
    }
    if gather_ms is not None:
        line["final_gather_ms"] = gather_ms
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
