set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
AB_ROUNDS=2 timeout 600 python tools/split_ab.py in-tree > gpurun_out/r2_split_ab6.txt 2>&1
AB_MODEL=gsc12 AB_ROUNDS=1 timeout 600 python tools/split_ab.py in-tree:0 in-tree >> gpurun_out/r2_split_ab6.txt 2>&1
cat gpurun_out/r2_split_ab6.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split" > gpurun_out/r2_pytest_split6.txt 2>&1
tail -5 gpurun_out/r2_pytest_split6.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cepstral" -s 1 -c 1 -o gpurun_out/r2_split6 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_split6.log 2>&1
tail -3 gpurun_out/r2_ncu_split6.log
