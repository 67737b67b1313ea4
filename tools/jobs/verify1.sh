set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_n.txt 2>&1
tail -5 gpurun_out/r2_pytest_n.txt
timeout 900 python bench.py > gpurun_out/r2_bench_n.json 2> gpurun_out/r2_bench_n.err
tail -c 600 gpurun_out/r2_bench_n.json; tail -3 gpurun_out/r2_bench_n.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_n_ref.json 2>> gpurun_out/r2_bench_n.err
tail -c 400 gpurun_out/r2_bench_n_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu_n.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r2_sanitizer_memcheck_n.txt 2>&1
tail -3 gpurun_out/r2_sanitizer_memcheck_n.txt
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_racecheck_n.txt 2>&1
tail -3 gpurun_out/r2_sanitizer_racecheck_n.txt
