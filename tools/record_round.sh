#!/bin/bash
# Produces the raw material of a round's measurement record on the GPU box (run under gpurun):
#   gpurun --timeout 1500 -- 'bash tools/record_round.sh r2p'
# Everything lands in gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into the JSON kept under profiles/.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest_gpu.txt
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_err.log
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2>> $out/${tag}_err.log
for w in c1 c3 c5s; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --also none > $out/${tag}_bench_$w.json 2>> $out/${tag}_err.log
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also none --no-graph > $out/${tag}_bench_nograph.json 2>> $out/${tag}_err.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also none --force-f64 > $out/${tag}_bench_f64.json 2>> $out/${tag}_err.log
for m in host whole; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also none --ri-upload $m > $out/${tag}_bench_upload_$m.json 2>> $out/${tag}_err.log
done
# launch list of the bench command (cold-cache, serialised: shares of the step, not absolute times)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --also none --no-graph > $out/${tag}_ncu_list.log 2>&1
# full capture of every kernel of one step (direct launches; the first steps -- set-up, warm-up -- are skipped)
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 40 -c 11 -f -o $out/${tag}_step \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --also none --no-graph > $out/${tag}_ncu_full.log 2>&1
# the operators of config 4: one full capture of each kernel of one DynamicScatter / Voxelization / scatter_v2 call
# (no source import here: the merged gpurun_out/ is limited to 64 MiB per call)
if [ "$2" = "ops" ]; then
ncu --set full --clock-control none -k regex:"k_|DeviceRadixSort|DeviceScan" -s 60 -c 24 -f -o $out/${tag}_ops \
    python tools/bench_ops.py > $out/${tag}_ncu_ops.log 2>&1
fi
python tools/bench_ops.py > $out/${tag}_ops_c4.jsonl 2>> $out/${tag}_err.log
cat $out/${tag}_pytest_gpu.txt
tail -2 $out/${tag}_err.log
