#!/bin/bash
# A/B of --batch-segments on a scaled c5 job (one GPU):  gpurun -- 'bash tools/experiments/job_batch_ab.sh 2560 "8 0"'
n=${1:-2560}
mkdir -p gpurun_out
for bs in ${2:-8 0}; do
  python bench.py --workload c5 --job-tracklets $n --steps 5 --warmup 2 --no-cpu-baseline --batch-segments $bs \
      > gpurun_out/jb_${n}_$bs.json 2> gpurun_out/jb_${n}_$bs.err
  python - $n $bs <<'PY'
import json, sys
n, bs = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/jb_{n}_{bs}.json")); j = d["job"]
    print(f"batch-segments {bs}: job {d['ms_per_step']:.3f} ms  value {d['value']:.0f}  batches {j['batches_per_rank']} x {j['batch_segments']}  frac {d['roofline']['frac']:.3f}  e2e {d['e2e']['value']:.0f}  mism {j['label_mismatches_vs_cpu_port']}")
except Exception as e:
    print("batch-segments", bs, "ERR", e, open(f"gpurun_out/jb_{n}_{bs}.err").read()[-400:])
PY
done
