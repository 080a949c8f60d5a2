"""Debug build only (-DOCC_VISDEBUG=1): per-warp timeline of k_visibility on one workload.
    OCCB200_LIB=.../libocc_b200_dbg.so python tools/experiments/vis_debug.py c2"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from objectcentricocccompletion_b200 import _lib, occ_annotate, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
batch = synth.config_batch(name, seed=0)
pk = occ_annotate.pack_tracklets(batch)
host = occ_annotate.HostBuffers(pk, pin=True)
d = occ_annotate.DeviceTracklets(pk, labels="u8")
d.upload(host)
for _ in range(3):
    d.run(0)
torch.cuda.synchronize()
L = C.CDLL(_lib.LIB_PATH)
n = 8 * 148 * 8 * 8
buf = np.zeros(n, np.int64)
assert L.occb200_debug_visibility(C.c_void_p(buf.ctypes.data), n) == 0
w = buf.reshape(-1, 8)
w = w[w[:, 1] > 0]
t0 = w[:, 0].min()
start, end = w[:, 0] - t0, w[:, 1] - t0
print("warps", len(w), "kernel span cycles", end.max(), "mean end", end.mean(), "median end", np.median(end))
print("start: min/median/max", start.min(), np.median(start), start.max())
print("items/warp mean", w[:, 2].mean(), "max", w[:, 2].max(), " iterations/warp mean", w[:, 3].mean(), "max", w[:, 3].max(), "total", w[:, 3].sum())
busy = end - start
print("busy cycles/warp mean", busy.mean(), "max", busy.max(), " cycles per iteration (mean busy/mean iters)", busy.mean() / max(w[:, 3].mean(), 1))
print("longest single item cycles: mean", w[:, 4].mean(), "max", w[:, 4].max(), "slice of the max", np.bincount(w[:, 5].astype(int)).tolist())
q = np.quantile(end, [0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 1.0])
print("end-time quantiles", q.astype(int).tolist())
late = w[end > 0.8 * end.max()]
print("warps ending in the last 20% of the span:", len(late), " their iterations mean", late[:, 3].mean(), "items", late[:, 2].mean(),
      "longest item mean", late[:, 4].mean())
sm = w[:, 6]
per_sm_end = np.array([end[sm == s].max() for s in np.unique(sm)])
print("per-SM last end: min", per_sm_end.min(), "mean", per_sm_end.mean(), "max", per_sm_end.max())
