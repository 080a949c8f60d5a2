#!/usr/bin/env python
"""Summarise an ncu report (``ncu --set full``) into the small JSON kept under ``profiles/``: per launch the
duration, DRAM bytes, issue utilisation, occupancy, local-memory traffic and the top stall reasons; at the top
level the sums over the launches whose kernel name matches ``--sum`` (default: the ray-cast kernels, which
``bench.py`` reads for ``roofline.traffic``).

    python tools/ncu_summary.py gpurun_out/step.ncu-rep profiles/r2_step_ncu.json "capture command line" [--sum regex]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
         "nsecond": 1e-3, "msecond": 1e3}


def main():
    import re
    argv = list(sys.argv)
    pat = "k_brick_cull|k_visibility\\("
    if "--sum" in argv:
        i = argv.index("--sum")
        pat = argv[i + 1]
        del argv[i:i + 2]
    rep, out, capture = argv[1], argv[2], (argv[3] if len(argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units = rows[0], rows[1]
    launches, total = [], {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    for r in rows[2:]:
        rec = dict(zip(names, r))
        m = {}
        for k in KEEP:
            if k in rec and rec[k] != "":
                m[k] = {"unit": units[names.index(k)], "value": rec[k]}
        stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
                  for k, v in rec.items() if k.startswith("smsp__average_warps_issue_stalled_")
                  and k.endswith("_per_issue_active.ratio") and v not in ("", None)}
        top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        name = rec.get("Kernel Name", "")
        launches.append({"kernel": name[:80], "metrics": m, "stall_cycles_per_issue": top})
    # the sums take ONE launch per distinct kernel name, the last one captured (a capture window may hold the same kernel
    # of two consecutive steps; "per step" must not count it twice)
    last = {}
    for L in launches:
        if re.search(pat, L["kernel"]):
            last[L["kernel"].split("(")[0]] = L["metrics"]
    for m in last.values():
        for k in total:
            if k in m:
                total[k] += float(m[k]["value"]) * SCALE.get(m[k]["unit"], 1.0)
    doc = {
        "kernel": "the last captured launch of every kernel matching /%s/ (sums below)" % pat,
        "capture": capture,
        "metrics": {
            "gpu__time_duration.sum": {"unit": "us", "value": "%.3f" % total["gpu__time_duration.sum"]},
            "dram__bytes_read.sum": {"unit": "byte", "value": "%.0f" % total["dram__bytes_read.sum"]},
            "dram__bytes_write.sum": {"unit": "byte", "value": "%.0f" % total["dram__bytes_write.sum"]},
        },
        "launches": launches,
    }
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc["metrics"]))


if __name__ == "__main__":
    main()
