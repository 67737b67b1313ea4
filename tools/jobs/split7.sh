set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or golden or float" > gpurun_out/r2_pytest_split7.txt 2>&1
tail -8 gpurun_out/r2_pytest_split7.txt
timeout 600 python bench.py --model l476f32 --f32-input --clips-per-gpu 262144 --steps 5 --no-cpu-baseline --no-also > gpurun_out/r2_bench_l_f32.json 2> gpurun_out/r2_bench_l_f32.err
tail -c 2500 gpurun_out/r2_bench_l_f32.json; tail -3 gpurun_out/r2_bench_l_f32.err
