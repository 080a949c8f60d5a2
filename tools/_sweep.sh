mkdir -p gpurun_out
for p in A B C D E F G; do
  OCCB200_LIB=$PWD/objectcentricocccompletion_b200/csrc/_build/libvar_$p.so timeout 300 python -m pytest tests/test_annotate_gpu.py -m gpu -x -q -k "vs_oracle or edge" 2>&1 | tail -1
  for w in c2 c5s; do
    OCCB200_LIB=$PWD/objectcentricocccompletion_b200/csrc/_build/libvar_$p.so timeout 200 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/vx_${w}_$p.json 2>>gpurun_out/sw_err.log
  done
done
tail -3 gpurun_out/sw_err.log
