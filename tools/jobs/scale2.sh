set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q -m gpu -k "multi or shards or every_gpu" > gpurun_out/r2_pytest_2gpu.txt 2>&1
tail -3 gpurun_out/r2_pytest_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-also > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n2_final.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['frac_of_ceiling'])"
