set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split or golden or float" > gpurun_out/r2_pytest_split9.txt 2>&1
tail -4 gpurun_out/r2_pytest_split9.txt
timeout 600 python bench.py --model l476f32 --f32-input --clips-per-gpu 262144 --steps 5 --no-cpu-baseline --no-also --e2e-steps 1 > gpurun_out/r2_bench_m_f32.json 2> gpurun_out/r2_bench_m_f32.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_m_f32.json').read().strip().splitlines()[-1]); print('6 CTAs', d['value'], d['roofline']['kernels'])"
EIKWS_B200_LIB=$PWD/ab/libeikws_f5.so timeout 600 python bench.py --model l476f32 --f32-input --clips-per-gpu 262144 --steps 5 --no-cpu-baseline --no-also --e2e-steps 1 > gpurun_out/r2_bench_m_f32_5ctas.json 2>> gpurun_out/r2_bench_m_f32.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_m_f32_5ctas.json').read().strip().splitlines()[-1]); print('5 CTAs', d['value'], d['roofline']['kernels'])"
EIKWS_MODEL=l476f32 EIKWS_F32=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cepstral" -s 1 -c 1 -o gpurun_out/r2_split9 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_split9.log 2>&1
tail -2 gpurun_out/r2_ncu_split9.log
