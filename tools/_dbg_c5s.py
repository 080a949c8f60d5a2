"""Debug: determinism of the annotate pipeline at c5s scale, single-stream and two-pipe."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from objectcentricocccompletion_b200 import occ_annotate, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "c5s"
dev = torch.device("cuda:0")
batch = synth.config_batch(wl)
pk = occ_annotate.pack_tracklets(batch)
host = occ_annotate.HostBuffers(pk)
KEYS = ("labels", "dims", "status", "n_unknown")


def snap(d):
    torch.cuda.synchronize()
    return {k: getattr(d, k).cpu().numpy().copy() for k in KEYS}


def diff(a, b, tag):
    bad = False
    for k in KEYS:
        ne = a[k] != b[k]
        if ne.any():
            bad = True
            idx = np.flatnonzero(ne.reshape(-1))
            print(f"  [{tag}] {k}: {idx.size} differ; first {idx[:8]}", flush=True)
            if k == "labels":
                t = np.searchsorted(pk.label_off, idx, side="right") - 1
                ts, cnt = np.unique(t, return_counts=True)
                print(f"    tracklets {ts[:16]} counts {cnt[:16]} (of {ts.size}); a={a[k].reshape(-1)[idx[:8]]} b={b[k].reshape(-1)[idx[:8]]}")
    if not bad:
        print(f"  [{tag}] identical", flush=True)
    return bad


for flags in (0, 2, 1):
    print("flags", flags, flush=True)
    d = occ_annotate.DeviceTracklets(pk, dev)
    d2 = occ_annotate.DeviceTracklets(pk, dev)
    d.upload(host); d2.upload(host)
    d.run(flags); r0 = snap(d)
    d.run(flags); r1 = snap(d)
    diff(r0, r1, "same buffer, run twice")
    d2.run(flags); r2 = snap(d2)
    diff(r0, r2, "second buffer, serial")
    if flags == 0:
        ref = r0
    else:
        diff(ref, r0, "vs flags 0")
    # two pipes
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for rep in range(3):
        for i in range(6):
            st, dd = (s1, d) if i % 2 == 0 else (s2, d2)
            with torch.cuda.stream(st):
                dd.upload(host)
                dd.run(flags)
        ra, rb = snap(d), snap(d2)
        diff(r0, ra, f"pipe0 rep{rep}")
        diff(r0, rb, f"pipe1 rep{rep}")
    # two pipes without upload
    for rep in range(2):
        for i in range(6):
            st, dd = (s1, d) if i % 2 == 0 else (s2, d2)
            with torch.cuda.stream(st):
                dd.run(flags)
        ra, rb = snap(d), snap(d2)
        diff(r0, ra, f"noupload pipe0 rep{rep}")
        diff(r0, rb, f"noupload pipe1 rep{rep}")
    del d, d2
