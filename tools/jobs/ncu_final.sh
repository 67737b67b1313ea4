set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel|cepstral" -s 2 -c 2 -o gpurun_out/r2_final_i16 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_final_i16.log 2>&1
tail -2 gpurun_out/r2_ncu_final_i16.log
EIKWS_MODEL=l476f32 EIKWS_F32=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel|cepstral" -s 2 -c 2 -o gpurun_out/r2_final_f32 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_final_f32.log 2>&1
tail -2 gpurun_out/r2_ncu_final_f32.log
