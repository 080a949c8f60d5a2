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
g0 = w[:, 7].min()
start, end = (w[:, 7] - g0) / 1e3, (w[:, 1] - g0) / 1e3            # microseconds on the global timer
busy_cyc = w[:, 0]
print("warps", len(w), "kernel span us", round(end.max(), 2))
print("start us: quantiles 0/10/50/90/100", np.quantile(start, [0, .1, .5, .9, 1]).round(2).tolist())
print("end   us: quantiles 0/10/25/50/75/90/100", np.quantile(end, [0, .1, .25, .5, .75, .9, 1]).round(2).tolist())
print("items/warp mean", w[:, 2].mean().round(2), "max", w[:, 2].max(), " iterations/warp mean", w[:, 3].mean().round(2), "max", w[:, 3].max(), "total", w[:, 3].sum())
print("busy cycles/warp mean", busy_cyc.mean().round(0), "max", busy_cyc.max(), " cycles per iteration", (busy_cyc.mean() / max(w[:, 3].mean(), 1)).round(0))
print("longest single item cycles: mean", w[:, 4].mean().round(0), "max", w[:, 4].max())
# how many warps are still running at time t
for t in np.linspace(0, end.max(), 13):
    print("  t=%6.2f us  running warps %5d" % (t, int(((start <= t) & (end > t)).sum())))
