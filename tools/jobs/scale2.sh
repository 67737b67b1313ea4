set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q -m gpu -k "multi or shards or every_gpu or split" > gpurun_out/r2_pytest_2gpu.txt 2>&1
tail -5 gpurun_out/r2_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 1500 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
