cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 300 python tools/kernel_lab.py --tag new --no-parity 2>&1 | tail -1
GAUDI_B200_LIB=$PWD/gaudi_b200/csrc/lib_prev.so timeout 300 python tools/kernel_lab.py --tag prev --no-parity 2>&1 | tail -1
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r2d.json
