set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
AB_ROUNDS=2 timeout 600 python tools/split_ab.py in-tree > gpurun_out/r2_split_ab12.txt 2>&1
cat gpurun_out/r2_split_ab12.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_o.txt 2>&1
tail -4 gpurun_out/r2_pytest_o.txt
timeout 900 python bench.py > gpurun_out/r2_bench_o.json 2> gpurun_out/r2_bench_o.err
tail -c 300 gpurun_out/r2_bench_o.json; tail -3 gpurun_out/r2_bench_o.err
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_memcheck_o.txt 2>&1
tail -2 gpurun_out/r2_sanitizer_memcheck_o.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_racecheck_o.txt 2>&1
tail -2 gpurun_out/r2_sanitizer_racecheck_o.txt
