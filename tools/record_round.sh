#!/bin/bash
# Produces the raw material of a round's measurement record on the GPU box (run under gpurun):
#   gpurun --timeout 1500 -- 'bash tools/record_round.sh r1q'
# Everything lands in gpurun_out/; the summaries worth keeping are copied into profiles/ by hand afterwards.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest_gpu.txt
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_err.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2>> $out/${tag}_err.log
for w in c1 c3 c5s; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$w.json 2>> $out/${tag}_err.log
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > $out/${tag}_bench_nograph.json 2>> $out/${tag}_err.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --force-f64 > $out/${tag}_bench_f64.json 2>> $out/${tag}_err.log
# launch list of the bench command (cold-cache, serialised: shares of the step, not absolute times)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1
# full capture of the ray-cast launches of one step (skip the warm-up / capture runs)
ncu --set full --clock-control none --import-source on -k regex:k_visibility_fast -s 6 -c 2 -f -o $out/${tag}_vis \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
python tools/bench_ops.py > $out/${tag}_ops_c4.jsonl 2>> $out/${tag}_err.log
cat $out/${tag}_pytest_gpu.txt
tail -2 $out/${tag}_err.log
