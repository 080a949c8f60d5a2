"""Where the one-shot annotate_batch() call spends its time (phases separated by device synchronisation).
    python tools/experiments/api_breakdown.py [c2]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from objectcentricocccompletion_b200 import occ_annotate, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
MODE = {"arena": "arena", "pageable": False}[sys.argv[2] if len(sys.argv) > 2 else "arena"]
batch = synth.config_batch(name, seed=0)
for rep in range(4):
    t = [time.perf_counter()]

    def lap():
        torch.cuda.synchronize()
        t.append(time.perf_counter())

    pk = occ_annotate.pack_tracklets(batch); lap()
    host = occ_annotate.HostBuffers(pk, pin=MODE, windows=True); lap()
    dev = occ_annotate.DeviceTracklets(pk); lap()
    dev.upload(host); lap()
    dev.run(0); lap()
    res = dev.results(); lap()
    names = ["pack", "HostBuffers", "DeviceTracklets", "upload", "run", "results"]
    d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
    print("rep", rep, " ".join(f"{n} {x:.2f}" for n, x in zip(names, d)), "total %.2f ms" % sum(d), "bytes", host.nbytes())
    t0 = time.perf_counter()
    occ_annotate.annotate_batch(batch)
    torch.cuda.synchronize()
    print("   annotate_batch %.2f ms" % (1e3 * (time.perf_counter() - t0)))
