#!/bin/bash
# A/B of --job-streams on a scaled c5 job (one GPU):  gpurun -- 'bash tools/experiments/job_streams_ab.sh 2560 "1 2 3"'
n=${1:-2560}
mkdir -p gpurun_out
for js in ${2:-1 2}; do
  python bench.py --workload c5 --job-tracklets $n --steps 5 --warmup 2 --no-cpu-baseline --job-streams $js \
      > gpurun_out/js_${n}_$js.json 2> gpurun_out/js_${n}_$js.err
  python - $n $js <<'PY'
import json, sys
n, js = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/js_{n}_{js}.json"))
    j = d["job"]
    print(f"streams {js}: job {d['ms_per_step']:.3f} ms  value {d['value']:.0f}  compute {j['compute_ms_max_over_ranks']:.3f}  gather {j['final_gather_ms']:.3f}  frac {d['roofline']['frac']:.3f}  e2e {d['e2e']['value']:.0f}  mism {j['label_mismatches_vs_cpu_port']}")
except Exception as e:
    print("streams", js, "ERR", e, open(f"gpurun_out/js_{n}_{js}.err").read()[-400:])
PY
done
