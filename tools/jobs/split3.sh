set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tools/split_ab.py in-tree:0 in-tree ab/libeikws_tma.so ab/libeikws_tma12.so ab/libeikws_c6.so > gpurun_out/r2_split_ab3.txt 2>&1
cat gpurun_out/r2_split_ab3.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_j.txt 2>&1
tail -5 gpurun_out/r2_pytest_j.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel|cepstral" -s 2 -c 2 -o gpurun_out/r2_split3 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_split3.log 2>&1
tail -3 gpurun_out/r2_ncu_split3.log
