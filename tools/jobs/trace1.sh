set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
EIKWS_B200_LIB=$PWD/ab/libeikws_trace.so timeout 300 python tools/cep_trace.py > gpurun_out/r2_cep_trace.txt 2>&1
cat gpurun_out/r2_cep_trace.txt
