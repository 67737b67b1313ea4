set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
AB_ROUNDS=2 timeout 600 python tools/split_ab.py in-tree ab/libeikws_cd1.so ab/libeikws_cd2.so ab/libeikws_cd3.so > gpurun_out/r2_split_ab4.txt 2>&1
cat gpurun_out/r2_split_ab4.txt
timeout 900 python bench.py > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err
tail -c 3000 gpurun_out/r2_bench_k.json; tail -5 gpurun_out/r2_bench_k.err
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_memcheck_split.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_memcheck_split.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py --quick > gpurun_out/r2_sanitizer_racecheck_split.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_racecheck_split.txt
