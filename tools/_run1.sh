cd $GRAFT_REPO_ROOT
timeout 180 python tools/kernel_lab.py --tag tail --batch 512 --steps 2 2>&1 | tail -1
timeout 300 python tools/kernel_lab.py --tag tail --no-parity 2>&1 | tail -1
GAUDI_B200_LIB=$PWD/gaudi_b200/csrc/lib_prev.so timeout 300 python tools/kernel_lab.py --tag prev --no-parity 2>&1 | tail -1
timeout 300 python tools/kernel_lab.py --tag tail --no-parity 2>&1 | tail -1
GAUDI_B200_LIB=$PWD/gaudi_b200/csrc/lib_prev.so timeout 300 python tools/kernel_lab.py --tag prev --no-parity 2>&1 | tail -1
