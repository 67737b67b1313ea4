set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tools/split_ab.py in-tree:0 in-tree ab/libeikws_w12.so ab/libeikws_w14.so ab/libeikws_c4.so ab/libeikws_c6.so > gpurun_out/r2_split_ab2.txt 2>&1
cat gpurun_out/r2_split_ab2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split" > gpurun_out/r2_pytest_split2.txt 2>&1
tail -5 gpurun_out/r2_pytest_split2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cepstral" -s 1 -c 1 -o gpurun_out/r2_split2 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_split2.log 2>&1
tail -3 gpurun_out/r2_ncu_split2.log
