#!/bin/bash
# A/B of library variants built by tools/build_variants.sh, on the GPU box:
#   gpurun -- 'bash tools/ab.sh tag "c2 c5s" variant1 variant2 ...'
# Per variant: the annotate parity tests, then bench.py per workload (no CPU leg); one summary line each.
tag=$1; shift
wls=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=$PWD/objectcentricocccompletion_b200/csrc/_build/variants/libocc_b200_$v.so
  [ "$v" = "main" ] && lib=$PWD/objectcentricocccompletion_b200/csrc/libocc_b200.so
  OCCB200_LIB=$lib python -m pytest tests/test_annotate_gpu.py -x -q -k "vs_oracle or fixture or edge" 2>&1 | tail -1 > gpurun_out/${tag}_${v}_parity.txt
  for w in $wls; do
    OCCB200_LIB=$lib python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --also none \
        > gpurun_out/${tag}_${v}_$w.json 2> gpurun_out/${tag}_${v}_$w.err
  done
done
python - "$tag" "$wls" "$@" <<'PY'
import json, sys
tag, wls, vs = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
for v in vs:
    par = open(f"gpurun_out/{tag}_{v}_parity.txt").read().strip()
    for w in wls:
        try:
            d = json.load(open(f"gpurun_out/{tag}_{v}_{w}.json"))
            k = d["roofline"]["kernels_ms"]
            print(f"{v:10s} {w:4s} step {d['ms_per_step']:.4f} cull {k['k_brick_cull']:.4f} vis {k['k_visibility']:.4f} "
                  f"frac {d['roofline']['frac']:.3f} crop {k['k_crop_voxelize']:.4f} pb {k['k_pair_build']:.4f} "
                  f"exec {d['executed_steps_per_step']} e2e {d['e2e']['value']:.0f} | {par}")
        except Exception as e:
            print(v, w, "ERR", e, open(f"gpurun_out/{tag}_{v}_{w}.err").read()[-300:])
PY
