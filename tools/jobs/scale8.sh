set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_split.json 2> gpurun_out/r2_bench_n8_split.err
tail -c 600 gpurun_out/r2_bench_n8_split.json; tail -3 gpurun_out/r2_bench_n8_split.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 --no-also > gpurun_out/r2_bench_n4_split.json 2> gpurun_out/r2_bench_n4_split.err
tail -c 300 gpurun_out/r2_bench_n4_split.json
