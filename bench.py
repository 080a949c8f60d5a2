#!/usr/bin/env python
"""Benchmark of the point -> occupancy hot path (BASELINE.json metric: tracklets/s and voxel-steps/s
of the occupancy ray-cast, %HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference] [--also c1,c3,c4,c5]

A *step* is one pass of the whole annotate path (crop -> box frame -> voxelise -> range-image
visibility) over the workload's synthetic tracklets.

N = 1 (default): the workload is BASELINE.json configs[1] (C2): 64 vehicle tracklets x 40 frames at 0.2 m
voxels, one shared segment of 5-LiDAR range images.  The line also carries `also`: sub-records for C1, C3, C4
and the 10 000-tracklet job C5 on this one GPU (ms/step, roofline fraction, label mismatches against the CPU
port on a fixed subsample -- asserted, not just printed).

N > 1 (torchrun, one rank per GPU): STRONG scaling of the fixed job C5 = BASELINE.json configs[4]: 10 000
C2-shaped tracklets over 157 segments, sharded by segment (longest-processing-time greedy, dist.shard_indices),
every rank generates and annotates only its own segments in batches of `--batch-segments` segments, and the step
ends with the job's one collective: the gather of all labels (one byte per voxel, device to device) on rank 0.
`value` = 10 000 tracklets / max over ranks of {compute of all its batches + gather}.

`value`     tracklets/s with the inputs resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`       the same through the public API with HOST buffers: per step, H2D of all inputs from pinned
            memory + the kernels + D2H of labels/dims/status; steps alternate between two device buffer sets
            on two streams, so one step's upload overlaps the previous step's kernels (events around all K).
            `pack_ms` (host packing of the step's metadata; the large arrays are uploaded from where they lie)
            and `api_one_shot_ms` (occ.annotate_batch(): pack + allocate + upload from pageable memory + run +
            download) are reported beside it.
`roofline`  the ray-cast (k_brick_cull + k_visibility): algorithmic bytes 4*U*B*L + 4*V per tracklet
             / its mean launch duration, measured live with CUDA events
            recorded around the kernels on the launching stream (occb200_profile_*).
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
    "c5": "c5: 10000 vehicle tracklets x 40 frames over 157 segments, 0.2 m voxels, 5 LiDARs, sharded by segment",
}
JOB_TRACKLETS, JOB_PER_SEGMENT = 10000, 64


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 on one GPU, the c5 job on several")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-threads", type=int, default=0, help="0 = all host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-f64", action="store_true", help="all-f64 visibility kernel")
    ap.add_argument("--flags", type=int, default=0, help="extra occb200_annotate_args_t.flags bits (A/B measurements)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--also", default="c1,c3,c4,c5", help="N=1: sub-records to add to the line (comma list, 'none')")
    ap.add_argument("--batch-segments", type=int, default=0,
                    help="c5 job: segments per device call (0 = a rank's segments in equal batches of at most 16)")
    ap.add_argument("--job-tracklets", type=int, default=JOB_TRACKLETS, help="c5 job size (tests use a small one)")
    ap.add_argument("--job-streams", type=int, default=2, help="c5 job: streams the batches' graphs alternate on")
    ap.add_argument("--ri-upload", default="pull", choices=["pull", "host", "whole"],
                    help="e2e leg: how the range images reach the device (pull: the device fetches the windows it "
                         "can read from pinned host memory; host: the host gathers them; whole: every image)")
    return ap.parse_args(argv)


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
    """DRAM bytes per step of the ray-cast kernels from the COMMITTED ncu capture (profiles/): a static record of
    the build that was profiled, not a measurement of this run."""
    p = os.path.join(ROOT, "profiles", "raycast_ncu.json")
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
    ok = [r for r in res if r["occ"] is not None]
    U = sum(r["n_unknown"] for r in ok)
    V = sum(int(r["occ"].size) for r in ok)
    steps = U * B * L
    return dict(U=U, V=V, steps=steps, vis_bytes=4 * steps + 4 * V, executed=sum(r.get("n_steps", 0) for r in ok),
                ok=len(ok))


KERNEL_KINDS = ("k_crop_voxelize", "k_tracklet_setup (f64 path) / side: redo crop pass", "k_scan_chunks", "k_brick_cull", "k_visibility",
                "side:k_pyr_build+k_table_setup", "k_visibility_recheck", "k_pair_build", "k_labels")


def workload_config(name, batch, T=None):
    B, L = len(batch.tracklets[0]), len(batch.segments[0].inclinations)
    return {"workload": WORKLOADS[name], "tracklets_per_step": int(T if T is not None else len(batch.tracklets)),
            "frames": B, "lidars": L, "voxel_size": batch.voxel_size}


def full_config(name, T, B, L, voxel_size, world, args):
    """The config both arms print (the reference arm echoes the CUDA arm's keys so that the two lines describe the
    same workload key for key)."""
    cfg = {"workload": WORKLOADS[name], "tracklets_per_step": int(T), "frames": int(B), "lidars": int(L),
           "voxel_size": float(voxel_size)}
    if name == "c5":
        cfg.update({"segments": (int(T) + JOB_PER_SEGMENT - 1) // JOB_PER_SEGMENT,
                    "sharding": "by segment, LPT (dist.shard_indices)",
                    "batch_segments": args.batch_segments if args.batch_segments > 0 else "auto (equal batches of <= 16 segments per rank)",
                    "final_gather": "inside the timed step (uint8 labels, device to device)",
                    "l2": "not flushed: every rank's inputs per step exceed L2 many times over",
                    "launch": "kernel by kernel" if args.no_graph else
                    f"cuda graph replay per batch, batches alternating on {max(1, args.job_streams)} stream(s)"})
    else:
        cfg.update({"l2": "flushed (256 MiB write) between timed steps",
                    "visibility": "f64" if args.force_f64 else "default",
                    "labels": "uint8 on the device (int32 at the file boundary)",
                    "launch": "kernel by kernel" if args.no_graph else "cuda graph replay",
                    "parallelism": "one batch of this shape per GPU" if world > 1 else "1 GPU"})
    return cfg


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


def count_mismatches(got, exp):
    """Labels / statuses that differ between two result lists (0 = parity)."""
    bad = 0
    for g, e in zip(got, exp):
        if g["status"] != e["status"]:
            bad += 1
        elif e["occ"] is not None:
            bad += 1 if g["occ"].shape != e["occ"].shape else int((g["occ"] != e["occ"]).sum())
    return bad


def job_sample(n_tracklets, seed=0):
    """The bounded sample of the c5 job the CPU arm times: its first two segments (128 tracklets)."""
    from objectcentricocccompletion_b200 import synth

    nseg = (n_tracklets + JOB_PER_SEGMENT - 1) // JOB_PER_SEGMENT
    return synth.make_batch(n_tracklets, 40, 0.2, "vehicle", seed, tracklets_per_segment=JOB_PER_SEGMENT,
                            only_segments=range(min(2, nseg)))


def run_reference(args):
    """--impl reference: the CPU port on all host cores, same config/metric/unit; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    os.environ["OCCB200_HOST_ONLY"] = "1"               # the package then loads no CUDA library: this arm is the CPU port only
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle

    name = args.workload or ("c2" if args.gpus == 1 else "c5")
    threads = args.cpu_threads or os.cpu_count()
    if name == "c5":
        batch = job_sample(args.job_tracklets)
        T_cfg, sample = args.job_tracklets, (f"first {len(batch.segments)} segments of the job ({len(batch.tracklets)} "
                                             f"tracklets) per step; the job is {args.job_tracklets} such tracklets")
    else:
        batch = synth.config_batch(name, seed=0)
        T_cfg, sample = len(batch.tracklets), f"full {name} batch ({len(batch.tracklets)} tracklets) per step"
    B, L = len(batch.tracklets[0]), len(batch.segments[0].inclinations)
    pk = oracle.PackedBatch(batch)
    for _ in range(args.warmup):
        oracle.annotate_batch(batch, threads=threads, packed=pk)
    steps = max(args.steps, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = oracle.annotate_batch(batch, threads=threads, packed=pk)
    dt = (time.perf_counter() - t0) / steps
    ws = workload_stats(res, B, L)
    T = len(batch.tracklets)
    val = T / dt
    cfg = full_config(name, T_cfg, B, L, batch.voxel_size, args.gpus, args)
    line = {"impl": "reference", "metric": "tracklets_per_s", "value": val, "unit": "tracklets/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * (T_cfg / T), "higher_is_better": True,
            "scaling": "weak" if name != "c5" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "ok_tracklets_in_sample": ws["ok"], "voxel_steps_per_s": ws["steps"] / dt,
            "cpu_baseline": {"value": val, "unit": "tracklets/s", "cores": threads, "kind": "port",
                             "sample": sample + f", {steps} steps, OpenMP over tracklets"},
            "e2e": {"value": val, "unit": "tracklets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class Ctx:
    """torch / device / distributed handles shared by the measurement routines."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.args = args
        self.flags = (1 if args.force_f64 else 0) | args.flags
        self.use_graph = not args.no_graph

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, step_fn, n, flush=True):
        """n steps, one CUDA-event pair each (L2 evicted in between, outside the timed region); total ms."""
        torch = self.torch
        evs = []
        for _ in range(n):
            if flush:
                self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    def profile(self, step_fn, n, flush=True):
        """The same steps with events around each kernel (direct launches) -> (ms per kind, launches per kind)."""
        from objectcentricocccompletion_b200 import _lib

        _lib.lib().occb200_profile_enable(1)
        self.timed(step_fn, n, flush)
        _lib.lib().occb200_profile_enable(0)
        nk = _lib.lib().occb200_profile_kinds()
        kms, kn = np.zeros(nk, np.float64), np.zeros(nk, np.int64)
        _lib.check(_lib.lib().occb200_profile_read(kms.ctypes.data, kn.ctypes.data), "occb200_profile_read")
        return kms, kn


def roofline_record(ws, kms, n_steps, peak, peak_src, traffic=None):
    vis_ms = (kms[3] + kms[4]) / max(n_steps, 1)             # the ray-cast: k_brick_cull + k_visibility
    step_ms = kms.sum() / max(n_steps, 1)
    achieved = ws["vis_bytes"] / (vis_ms / 1e3) / 1e9 if vis_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": "k_brick_cull + k_visibility (the ray-cast)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": ("profiles/raycast_ncu.json (static: the committed ncu capture, not this run)"
                               if traffic else None),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": ws["vis_bytes"], "kernel_ms": vis_ms,
            "kernel_share_of_step": vis_ms / step_ms if step_ms else None,
            "kernels_ms": dict(zip(KERNEL_KINDS, (kms / max(n_steps, 1)).round(5).tolist()))}


def measure_resident(ctx, batch, warmup):
    """One batch resident in HBM, results downloaded once, the graph captured and warmed up."""
    from objectcentricocccompletion_b200 import occ_annotate

    torch = ctx.torch
    B, L = len(batch.tracklets[0]), len(batch.segments[0].inclinations)
    pk = occ_annotate.pack_tracklets(batch)                  # first call: lazy imports, torch's CPU thread pool ...
    t0 = time.perf_counter()
    pk = occ_annotate.pack_tracklets(batch)
    pack_ms = (time.perf_counter() - t0) * 1e3
    host = occ_annotate.HostBuffers(pk, pin=True, windows=ctx.args.ri_upload)
    d = occ_annotate.DeviceTracklets(pk, ctx.dev, labels="u8")
    d.upload(host)
    d.run(ctx.flags)
    torch.cuda.synchronize()
    res = d.results()
    n_recheck, _ = d.queue_stats(ctx.flags)
    ws = workload_stats(res, B, L)
    graph_kernels = d.capture(ctx.flags) if ctx.use_graph else 0
    step = (lambda: d.replay(ctx.flags)) if ctx.use_graph else (lambda: d.run(ctx.flags))
    for _ in range(warmup):
        step()
    return dict(pk=pk, host=host, d=d, res=res, ws=ws, n_recheck=n_recheck, pack_ms=pack_ms, step=step,
                graph_kernels=graph_kernels, B=B, L=L)


def also_annotate(ctx, name, peak, peak_src, oracle_subset):
    """Sub-record of another annotate workload: resident step time, roofline, parity on a subsample (asserted)."""
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle

    steps = max(3, min(ctx.args.steps, 10))
    batch = synth.config_batch(name, seed=0)
    m = measure_resident(ctx, batch, max(ctx.args.warmup, 3))
    ms = ctx.timed(m["step"], steps) / steps
    kms, kn = ctx.profile(lambda: m["d"].run(ctx.flags), steps)
    T = len(batch.tracklets)
    sel = list(range(T))[::max(1, T // oracle_subset)][:oracle_subset]
    sub = synth.TrackletBatch(segments=batch.segments, tracklets=[batch.tracklets[i] for i in sel],
                              voxel_size=batch.voxel_size)
    exp = oracle.annotate_batch(sub, threads=os.cpu_count())
    mism = count_mismatches([m["res"][i] for i in sel], exp)
    assert mism == 0, f"{name}: {mism} label mismatches against the CPU port"
    rl = roofline_record(m["ws"], kms, steps, peak, peak_src)
    return {"config": workload_config(name, batch), "ms_per_step": ms, "tracklets_per_s": T / (ms / 1e3),
            "voxel_steps_per_s": m["ws"]["steps"] / (ms / 1e3), "executed_steps_per_step": m["ws"]["executed"],
            "voxel_steps_per_step": m["ws"]["steps"], "f64_rechecks_per_step": m["n_recheck"],
            "roofline_frac": rl["frac"], "raycast_ms": rl["kernel_ms"], "kernels_ms": rl["kernels_ms"],
            "label_mismatches_vs_cpu_port": mism, "parity_sample": f"{len(sel)} of {T} tracklets"}


def also_ops(ctx, peak):
    """C4: OcCo-Net input build -- dynamic Voxelization + DynamicScatter over 32 tracklets x 32 frames x 1024 points
    (N = 1 048 576), device-timed with L2 flushed, algorithmic bytes of SURVEY.md section 8d, parity against the
    CPU oracle asserted, the reference's own CUDA kernels (compiled unmodified for sm_100a) beside them if built."""
    import objectcentricocccompletion_b200 as occ
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle

    torch, dev = ctx.torch, ctx.dev
    pts, bidx = synth.scatter_inputs(32, 32, 1024, 5, seed=0, full=True)
    N = pts.shape[0]
    p = torch.from_numpy(pts).to(dev)
    b = torch.from_numpy(bidx).to(dev)
    vs, pcr = [0.2, 0.2, 0.2], [-204.8, -204.8, -4, 204.8, 204.8, 8]

    def med(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        evs = []
        for _ in range(iters):
            ctx.flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            e.record()
            evs.append((a, e))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(e) for a, e in evs]))

    try:
        from oracle import build as obuild
        rv = obuild.load_ref("ref_voxel_layer_cuda")
    except Exception:
        rv = None
    out = {"N": int(N)}
    vox = occ.Voxelization(vs, pcr, -1)
    coors = vox(p)
    assert (coors.cpu().numpy() == oracle.dynamic_voxelize(pts, vs, pcr)).all(), "c4: voxelize mismatch vs CPU oracle"
    ms = med(lambda: vox(p))
    alg = 4 * 5 * N + 12 * N
    rec = {"ms": ms, "alg_bytes": alg, "frac": alg / ms / 1e6 / peak}
    if rv is not None:
        rc = torch.zeros((N, 3), dtype=torch.int32, device=dev)
        rec["reference_cuda_kernel_sm100a_ms"] = med(lambda: rv.dynamic_voxelize(p, rc, vs, pcr, 3), 5, 1)
    out["Voxelization(dynamic) C=5"] = rec
    coors4 = torch.cat([b[:, None].int(), coors], 1).contiguous()
    for C, mode in [(3, "mean"), (5, "mean"), (128, "max")]:
        f = (p[:, :C].contiguous() if C <= 5 else torch.randn(N, C, device=dev))
        ds = occ.DynamicScatter(vs, pcr, mode == "mean")
        vf, vc = ds(f, coors4)
        M = int(vf.shape[0])
        if C == 3:
            evf, evc = oracle.dynamic_scatter_batched(f.cpu().numpy(), coors4.cpu().numpy(), "mean")[:2]
            assert (vc.cpu().numpy() == evc).all(), "c4: scatter coords mismatch vs CPU oracle"
            assert np.allclose(vf.cpu().numpy(), evf, rtol=1e-5, atol=1e-6), "c4: scatter feats mismatch vs CPU oracle"
        ms = med(lambda: ds(f, coors4))
        alg = 4 * C * N + 16 * N + (4 * C + 16) * M
        rec = {"ms": ms, "M": M, "alg_bytes": alg, "frac": alg / ms / 1e6 / peak}
        if rv is not None:
            c3 = coors.contiguous()
            rec["ms_3col"] = med(lambda: occ.dynamic_scatter(f, c3, mode))
            rec["reference_cuda_kernel_sm100a_3col_ms"] = med(lambda: rv.dynamic_point_to_voxel_forward(f, c3, mode), 3, 1)
        out[f"DynamicScatter({mode}) C={C}"] = rec
        c64 = coors4.long()
        nf, nc, inv = occ.scatter_v2(f, c64, mode)
        out[f"scatter_v2({mode}) C={C}"] = {
            "ms": med(lambda: occ.scatter_v2(f, c64, mode)),
            "ms_reusing_unq_inv": med(lambda: occ.scatter_v2(f, c64, mode, unq_inv=inv, new_coors=nc))}
    out["config"] = {"workload": "c4: dynamic Voxelization + DynamicScatter mean/max over 32 tracklets x 32 frames x 1024 points",
                     "voxel_size": vs, "point_cloud_range": pcr, "l2": "flushed between iterations"}
    return out


# ------------------------------------------------------------------------------------------------
def run_job(ctx, n_tracklets, steps, warmup, peak, peak_src, with_e2e=True, parity_tracklets=32):
    """The c5 job on ctx.world GPUs: segments sharded by LPT, per-rank batches of --batch-segments segments, the
    gather of all labels on rank 0 inside the timed step.  Returns the record on rank 0, None elsewhere."""
    from objectcentricocccompletion_b200 import dist as occ_dist
    from objectcentricocccompletion_b200 import occ_annotate, synth
    from oracle import oracle

    torch, dev, dist = ctx.torch, ctx.dev, ctx.dist
    rank, world = ctx.rank, ctx.world
    nseg = (n_tracklets + JOB_PER_SEGMENT - 1) // JOB_PER_SEGMENT
    seg_cost = [float(min(JOB_PER_SEGMENT, n_tracklets - i * JOB_PER_SEGMENT)) for i in range(nseg)]
    mine = occ_dist.shard_indices(seg_cost, world)[rank]
    synth.set_threads(max(1, (os.cpu_count() or 1) // world))
    t0 = time.perf_counter()
    full = synth.make_batch(n_tracklets, 40, 0.2, "vehicle", 0, tracklets_per_segment=JOB_PER_SEGMENT,
                            only_segments=mine)
    gen_s = time.perf_counter() - t0
    B, L = 40, len(full.segments[0].inclinations) if full.segments else 5
    S = ctx.args.batch_segments                               # batches of S segments
    if S <= 0:                                                # auto: as few batches of <= 16 segments as possible, equal sizes
        nb = max(1, -(-len(full.segments) // 16))
        S = max(1, -(-len(full.segments) // nb))
    batches = []
    for a in range(0, len(full.segments), S):
        trks = [synth.Tracklet(boxes=t.boxes, points=t.points, segment=t.segment - a, frame_ids=t.frame_ids,
                               kind=t.kind, flat=t.flat) for t in full.tracklets if a <= t.segment < a + S]
        batches.append(synth.TrackletBatch(segments=full.segments[a: a + S], tracklets=trks, voxel_size=0.2))
    t0 = time.perf_counter()
    pks = [occ_annotate.pack_tracklets(b) for b in batches]
    pack_ms = (time.perf_counter() - t0) * 1e3
    n_local = int(sum(pk.total_slots for pk in pks))
    labels_u8 = torch.zeros(max(n_local, 1), dtype=torch.uint8, device=dev)
    devs, hosts, off = [], [], 0
    for pk in pks:
        hosts.append(occ_annotate.HostBuffers(pk, pin=with_e2e, windows=ctx.args.ri_upload if with_e2e else True))
        devs.append(occ_annotate.DeviceTracklets(pk, dev, labels="u8",
                                                 labels_u8=labels_u8[off: off + max(pk.total_slots, 1)]))
        off += pk.total_slots
    for d, h in zip(devs, hosts):
        d.upload(h)
        d.run(ctx.flags)
    torch.cuda.synchronize()
    res = [r for d in devs for r in d.results()]
    n_recheck = sum(d.queue_stats(ctx.flags)[0] for d in devs)
    ws = workload_stats(res, B, L)
    # parity on a fixed subsample of this rank's tracklets against the CPU port (asserted)
    if parity_tracklets and batches:
        b0 = batches[0]
        sel = list(range(len(b0.tracklets)))[::max(1, len(b0.tracklets) // parity_tracklets)][:parity_tracklets]
        sub = synth.TrackletBatch(segments=b0.segments, tracklets=[b0.tracklets[i] for i in sel], voxel_size=0.2)
        exp = oracle.annotate_batch(sub, threads=max(1, (os.cpu_count() or 1) // world))
        mism = count_mismatches([res[i] for i in sel], exp)
        assert mism == 0, f"c5 rank {rank}: {mism} label mismatches against the CPU port"
    sizes = occ_dist.exchange_sizes(n_local)              # set-up: the payload sizes, once per job
    gathered = torch.empty(max(sum(sizes), 1), dtype=torch.uint8, device=dev) if rank == 0 else None
    if ctx.use_graph:
        for d in devs:
            d.capture(ctx.flags)

    # the batches of a rank are independent: their graphs alternate on --job-streams streams (fork from / join into the
    # current stream), so that the latency-bound ends of one batch overlap the wide kernels of the next
    n_js = max(1, min(ctx.args.job_streams, len(devs))) if ctx.use_graph else 1
    job_streams = [torch.cuda.Stream(dev) for _ in range(n_js)] if n_js > 1 else []

    def compute():
        if n_js <= 1:
            for d in devs:
                d.replay(ctx.flags) if ctx.use_graph else d.run(ctx.flags)
            return
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        for st in job_streams:
            st.wait_event(fork)
        for i, d in enumerate(devs):
            with torch.cuda.stream(job_streams[i % n_js]):
                d.replay(ctx.flags)
        for st in job_streams:
            join = torch.cuda.Event()
            join.record(st)
            cur.wait_event(join)

    def gather():
        occ_dist.gather_labels(labels_u8, sizes, 0, None, gathered)

    def job_step():
        compute()
        gather()

    for _ in range(max(warmup, 1)):
        job_step()
    ctx.barrier()
    ms = ctx.timed(job_step, steps, flush=False)
    ctx.barrier()
    ms_compute = ctx.timed(compute, steps, flush=False)
    ctx.barrier()
    ms_gather = ctx.timed(gather, steps, flush=False)
    ctx.barrier()
    if rank == 0 and world > 1:
        assert bool((gathered[: sizes[0]] == labels_u8[: sizes[0]]).all()), "gathered labels differ from rank 0's own"
    n_prof = max(1, min(steps, 5))
    kms, kn = ctx.profile(lambda: [d.run(ctx.flags) for d in devs], n_prof, flush=False)
    # e2e: per step every batch is uploaded from pinned host memory, annotated, its labels / dims / status downloaded
    ms_e2e, h2d, d2h = 0.0, 0, 0
    if with_e2e:
        keys = ("labels_u8", "dims", "status", "n_unknown")
        outs = [{k: torch.empty_like(getattr(d, k), device="cpu").pin_memory() for k in keys} for d in devs]
        streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

        def e2e_run(n):
            for d in devs:
                d.pulled.zero_()
            torch.cuda.synchronize()
            start = torch.cuda.Event(enable_timing=True)
            start.record()
            for st in streams:
                st.wait_event(start)
            k = 0
            for _ in range(n):
                for d, h, o in zip(devs, hosts, outs):
                    if h.ri_mode == "host" and k >= len(devs):      # (each batch has its own staging buffer)
                        h.gather_windows()
                    with torch.cuda.stream(streams[k % 2]):
                        d.upload(h)
                        d.replay(ctx.flags) if ctx.use_graph else d.run(ctx.flags)
                        for key, buf in o.items():
                            buf.copy_(getattr(d, key), non_blocking=True)
                    k += 1
            ends = []
            for st in streams:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                ends.append(e)
            torch.cuda.synchronize()
            return max(start.elapsed_time(e) for e in ends)

        e2e_run(1)
        ctx.barrier()
        ms_e2e = e2e_run(steps)
        ctx.barrier()
        h2d = sum(h.nbytes() for h in hosts) + sum(d.pulled_bytes() for d in devs) // max(steps, 1)
        d2h = sum(b.numel() * b.element_size() for o in outs for b in o.values())
    # ---- reduce over ranks
    t = torch.tensor([ms, ms_compute, ms_gather, ms_e2e], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(len(full.tracklets)), float(ws["steps"]), float(ws["executed"]), float(ws["vis_bytes"]),
                        float(h2d), float(d2h), float(n_recheck), float(ws["ok"])], dtype=torch.float64, device=dev)
    gm = torch.tensor([gen_s, pack_ms], dtype=torch.float64, device=dev)
    per_rank = torch.tensor([ms_compute / steps, float(len(full.tracklets)), (kms[3] + kms[4]) / n_prof,
                             float(ws["vis_bytes"])], dtype=torch.float64, device=dev)
    ranks = [per_rank.clone() for _ in range(world)]
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        dist.all_gather(ranks, per_rank)
    if rank != 0:
        return None
    ms, ms_compute, ms_gather, ms_e2e = (float(x) for x in t)
    T_all = float(tot[0])
    sec = ms / 1e3 / steps
    pr = np.array([r.cpu().numpy() for r in ranks])
    fr = [float(b / (m / 1e3) / 1e9 / peak) if m > 0 else 0.0 for m, b in zip(pr[:, 2], pr[:, 3])]
    rec = {
        "value": T_all / sec, "ms_per_step": sec * 1e3, "tracklets": int(T_all), "segments": nseg,
        "voxel_steps_per_s": float(tot[1]) / sec, "executed_steps_per_s": float(tot[2]) / sec,
        "voxel_steps_per_step": float(tot[1]), "f64_rechecks_per_step": float(tot[6]), "ok_tracklets": int(tot[7]),
        "compute_ms_max_over_ranks": ms_compute / steps, "final_gather_ms": ms_gather / steps,
        "gather_share_of_step": (ms_gather / steps) / (sec * 1e3), "gather_bytes": int(sum(sizes)),
        "per_rank": {"compute_ms": pr[:, 0].round(4).tolist(), "tracklets": pr[:, 1].astype(int).tolist(),
                     "raycast_ms": pr[:, 2].round(4).tolist(), "roofline_frac": [round(x, 4) for x in fr],
                     "load_imbalance": float(pr[:, 0].max() / max(pr[:, 0].mean(), 1e-9))},
        # whole job: all ranks' algorithmic ray-cast bytes / the slowest rank's ray-cast time / (N x peak)
        "roofline_frac": (float(pr[:, 3].sum() / (pr[:, 2].max() / 1e3) / 1e9 / peak / world) if pr[:, 2].max() > 0 else 0.0),
        "batches_per_rank": len(batches), "batch_segments": S, "job_streams": n_js, "generate_s_max": float(gm[0]), "pack_ms_max": float(gm[1]),
        "label_mismatches_vs_cpu_port": 0, "parity_sample": f"{parity_tracklets} tracklets per rank",
        "kernels_ms_rank0": dict(zip(KERNEL_KINDS, (kms / n_prof).round(5).tolist())),
        "graph_kernels_per_step_rank0": int(sum(getattr(d, "_graphs", {}).get(ctx.flags, (None, 0))[1] for d in devs)),
    }
    if with_e2e:
        sec_e2e = ms_e2e / 1e3 / steps
        rec["e2e"] = {"value": T_all / sec_e2e, "unit": "tracklets/s", "h2d_bytes_per_step": int(tot[4]),
                      "d2h_bytes_per_step": int(tot[5]), "ms_per_step": sec_e2e * 1e3}
    return rec


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import objectcentricocccompletion_b200 as occ
    from objectcentricocccompletion_b200 import _lib, occ_annotate, synth

    ctx = Ctx(args)
    torch, dev, rank, world = ctx.torch, ctx.dev, ctx.rank, ctx.world
    name = args.workload or ("c2" if world == 1 else "c5")
    peak, peak_src = peaks()
    sampler = ClockSampler(ctx.local)

    # ============================ the c5 job: strong scaling over the ranks ============================
    if name == "c5":
        if rank == 0:
            sampler.start()
        rec = run_job(ctx, args.job_tracklets, args.steps, args.warmup, peak, peak_src)
        if rank == 0:
            clocks = sampler.stop()
            cpu = None
            if not args.no_cpu_baseline:
                threads = args.cpu_threads or os.cpu_count()
                sample = job_sample(args.job_tracklets)
                dt, _, reps = cpu_port(sample, threads)
                cpu = {"value": len(sample.tracklets) / dt, "unit": "tracklets/s", "cores": threads, "kind": "port",
                       "sample": f"first {len(sample.segments)} segments of the job ({len(sample.tracklets)} tracklets), "
                                 f"best of {reps}, OpenMP over tracklets, C port of the reference path"}
            launches = int(rec["graph_kernels_per_step_rank0"] * args.steps * world) if ctx.use_graph else None
            line = {"metric": "tracklets_per_s", "value": rec["value"], "unit": "tracklets/s", "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": full_config("c5", rec["tracklets"], 40, 5, 0.2, world, args),
                    "voxel_steps_per_s": rec["voxel_steps_per_s"], "executed_steps_per_s": rec["executed_steps_per_s"],
                    "e2e": rec.get("e2e"), "gpu_launches": launches,
                    "roofline": {"bound": "hbm", "kernel": "k_brick_cull + k_visibility (the ray-cast)", "unit": "GB/s",
                                 "peak": peak, "peak_source": peak_src, "frac": rec["roofline_frac"],
                                 "achieved": rec["roofline_frac"] * peak, "traffic": None,
                                 "per_rank_frac": rec["per_rank"]["roofline_frac"]},
                    "job": rec, "cpu_baseline": cpu, "clocks": clocks, "final_gather_ms": rec["final_gather_ms"]}
            print(json.dumps(line))
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return

    # ============================ one batch per rank (N = 1: the headline C2 line) ============================
    batch = synth.config_batch(name, seed=rank)
    T = len(batch.tracklets)
    m = measure_resident(ctx, batch, args.warmup)
    d, host, pk, res, ws = m["d"], m["host"], m["pk"], m["res"], m["ws"]
    step_resident = m["step"]
    ctx.barrier()
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = ctx.timed(step_resident, args.steps)
    launches = m["graph_kernels"] * args.steps if ctx.use_graph else _lib.launch_count() - l0
    ctx.barrier()
    kms, kn = ctx.profile(lambda: d.run(ctx.flags), args.steps)

    # ---- e2e: host buffers -> H2D -> kernels -> D2H every step, double-buffered on two streams so the upload
    # of step k+1 overlaps the kernels / download of step k (inputs per step 134 MB > L2: no flush needed) -------
    out_keys = ("labels_u8", "dims", "status", "n_unknown")
    out_host = {k: torch.empty_like(getattr(d, k), device="cpu").pin_memory() for k in out_keys}
    d2 = occ_annotate.DeviceTracklets(pk, dev, labels="u8")
    if ctx.use_graph:
        d2.upload(host)
        d2.capture(ctx.flags)
    out_host2 = {k: torch.empty_like(h).pin_memory() for k, h in out_host.items()}
    pipes = [(torch.cuda.Stream(dev), d, out_host), (torch.cuda.Stream(dev), d2, out_host2)]
    staged = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_run(n):
        for _, dd, _ in pipes:
            dd.pulled.zero_()
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        for st, _, _ in pipes:
            st.wait_event(start)
        for i in range(n):
            st, dd, oh = pipes[i % 2]
            if host.ri_mode == "host":                  # the host's share of the step: re-read the window blocks
                if i >= 1:
                    staged[(i - 1) % 2].synchronize()   # (one staging buffer: the previous upload must have left it)
                host.gather_windows()
            with torch.cuda.stream(st):
                dd.upload(host)
                staged[i % 2].record(st)
                dd.replay(ctx.flags) if ctx.use_graph else dd.run(ctx.flags)
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
    ctx.barrier()
    ms_e2e = e2e_run(args.steps)
    ctx.barrier()
    pulled_per_step = sum(dd.pulled_bytes() for _, dd, _ in pipes) / max(args.steps, 1)
    assert all(bool((out_host2[k] == out_host[k]).all()) for k in out_host), "pipelined e2e results differ"
    # the e2e outputs against the resident run's (whole-pipeline parity of the upload mode in use)
    res_e2e = d.results()
    assert count_mismatches(res_e2e, res) == 0, "e2e labels differ from the resident run"
    # the public one-shot call, for scale: pack + allocate + upload from pageable memory + run + download
    api_ms = float("inf")
    for _ in range(3):                               # best of 3: the first call also pays allocator warm-up
        t0 = time.perf_counter()
        occ.annotate_batch(batch, flags=ctx.flags)
        api_ms = min(api_ms, (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(T), float(ws["steps"]), float(ws["executed"])], dtype=torch.float64, device=dev)
    if ctx.dist is not None:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
        ctx.dist.all_reduce(tot, op=ctx.dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    T_all, steps_all, exec_all = (float(x) for x in tot)
    if rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return

    sec = ms / 1e3 / args.steps
    sec_e2e = ms_e2e / 1e3 / args.steps
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = args.cpu_threads or os.cpu_count()
        dt, cres, reps = cpu_port(batch, threads)
        mism = count_mismatches(res, cres)
        assert mism == 0, f"{name}: {mism} label mismatches between the CUDA path and the CPU port"
        cpu = {"value": T / dt, "unit": "tracklets/s", "cores": threads, "kind": "port",
               "sample": f"full {name} batch ({T} tracklets), best of {reps}, OpenMP over tracklets, "
                         f"C port of the reference path; labels vs GPU: {mism} mismatches (asserted)",
               "voxel_steps_per_s": ws["steps"] / dt}
    cfg = full_config(name, T, m["B"], m["L"], batch.voxel_size, world, args)
    line = {
        "metric": "tracklets_per_s", "value": T_all / sec, "unit": "tracklets/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "voxel_steps_per_s": steps_all / sec, "executed_steps_per_s": exec_all / sec,
        "voxel_steps_per_step": ws["steps"], "executed_steps_per_step": ws["executed"],
        "f64_rechecks_per_step": m["n_recheck"], "unknown_voxels_per_step": ws["U"], "voxels_per_step": ws["V"],
        "ok_tracklets": ws["ok"],
        "e2e": {"value": T_all / sec_e2e, "unit": "tracklets/s", "h2d_bytes_per_step": int(host.nbytes() + pulled_per_step),
                "d2h_bytes_per_step": sum(h.numel() * h.element_size() for h in out_host.values()),
                "ms_per_step": sec_e2e * 1e3, "pack_ms": m["pack_ms"], "api_one_shot_ms": api_ms,
                "range_images": {"upload": host.ri_mode, "bytes_per_step": int(pulled_per_step) if host.ri_mode == "pull"
                                 else int(host.nbytes() - host.small.numel() - sum(t.numel() for _, t in host.pt_parts)),
                                 "whole_images_bytes": int(4 * pk.ri_len)}},
        "gpu_launches": int(launches),
        "roofline": roofline_record(ws, kms, args.steps, peak, peak_src, ncu_traffic() if name == "c2" else None),
        "cpu_baseline": cpu, "clocks": clocks,
    }
    # ---- the other configs, on this GPU (N = 1 only) ----
    also = [a for a in args.also.split(",") if a and a != "none"] if world == 1 and name == "c2" else []
    if also:
        del d2, m, d, host, pipes
        line["also"] = {}
        for a in also:
            try:
                if a in ("c1", "c3"):
                    line["also"][a] = also_annotate(ctx, a, peak, peak_src, oracle_subset=4 if a == "c3" else 1)
                elif a == "c4":
                    line["also"][a] = also_ops(ctx, peak)
                elif a == "c5":
                    line["also"][a] = run_job(ctx, args.job_tracklets, max(3, min(args.steps, 5)), 1, peak, peak_src)
            except AssertionError:
                raise                                   # a parity failure must fail the bench
            except Exception as e:                      # noqa: BLE001  (e.g. out of host memory for the 16 GB job)
                line["also"][a] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
