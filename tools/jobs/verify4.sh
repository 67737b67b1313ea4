set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
AB_ROUNDS=2 timeout 300 python tools/split_ab.py in-tree > gpurun_out/r2_split_ab14.txt 2>&1
cat gpurun_out/r2_split_ab14.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_q.txt 2>&1
tail -4 gpurun_out/r2_pytest_q.txt
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench_q.json 2> gpurun_out/r2_bench_q.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_q.json').read().strip().splitlines()[-1]); print(d['value'], [a['value'] for a in d['also']])"
