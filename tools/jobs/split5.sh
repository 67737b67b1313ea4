set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
tools/ubench/packed_mix > gpurun_out/r2_ubench_packed_mix.txt 2>&1
cat gpurun_out/r2_ubench_packed_mix.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel|cepstral" -s 2 -c 2 -o gpurun_out/r2_split5 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_split5.log 2>&1
tail -3 gpurun_out/r2_ncu_split5.log
