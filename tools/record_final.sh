python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/r2z_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_err.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2z_bench.json 2>> gpurun_out/r2z_err.log
for w in c1 c3 c5s; do python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --also none > gpurun_out/r2z_bench_$w.json 2>> gpurun_out/r2z_err.log; done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also none --ri-upload host > gpurun_out/r2z_bench_upload_host.json 2>> gpurun_out/r2z_err.log
python tools/bench_ops.py > gpurun_out/r2z_ops_c4.jsonl 2>> gpurun_out/r2z_err.log
cat gpurun_out/r2z_pytest_gpu.txt; tail -1 gpurun_out/r2z_smoke.txt; tail -2 gpurun_out/r2z_err.log
