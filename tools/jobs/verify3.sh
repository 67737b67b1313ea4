set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
AB_ROUNDS=2 timeout 600 python tools/split_ab.py in-tree ab/libeikws_statA.so > gpurun_out/r2_split_ab13.txt 2>&1
cat gpurun_out/r2_split_ab13.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_p.txt 2>&1
tail -4 gpurun_out/r2_pytest_p.txt
timeout 900 python bench.py > gpurun_out/r2_bench_p.json 2> gpurun_out/r2_bench_p.err
tail -c 300 gpurun_out/r2_bench_p.json; tail -3 gpurun_out/r2_bench_p.err
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_memcheck_p.txt 2>&1
tail -2 gpurun_out/r2_sanitizer_memcheck_p.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_racecheck_p.txt 2>&1
tail -2 gpurun_out/r2_sanitizer_racecheck_p.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_p.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu_p.log 2>&1
