set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/s5_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke.log 2>&1
python bench.py > gpurun_out/r01_bench_blj256.json 2> gpurun_out/s5_bench_blj256.err
python bench.py --workload lj38 > gpurun_out/r01_bench_lj38.json 2> gpurun_out/s5_bench_lj38.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_blj256.csv python bench.py --steps 2 --warmup 1 --pairs 8192 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:per_sf3_kernel -s 4 -c 2 -f -o gpurun_out/r01_prof_sf3_final python bench.py --steps 1 --warmup 1 --pairs 3256 --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/s5_tests.log gpurun_out/s5_smoke.log
