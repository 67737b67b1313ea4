set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python tools/split_ab.py 65536 10 > gpurun_out/r2_split_ab1.txt 2>&1
tail -20 gpurun_out/r2_split_ab1.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split" > gpurun_out/r2_pytest_split1.txt 2>&1
tail -15 gpurun_out/r2_pytest_split1.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_split_launches1.csv python tools/profile_run.py 65536 3 > gpurun_out/r2_split_launches1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"logmel|cepstral" -s 2 -c 2 -o gpurun_out/r2_split1 -f python tools/profile_run.py 16384 2 > gpurun_out/r2_ncu_split1.log 2>&1
tail -5 gpurun_out/r2_ncu_split1.log
