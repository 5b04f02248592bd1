timeout 60 python scripts/probe_ab.py blj256 2048 4 2>&1 | tail -1 | cut -c1-220
timeout 150 python -m pytest tests/test_periodic_gpu.py -m gpu -x -q 2>&1 | tail -3
FASTOVERLAP_B200_LIB=$PWD/fastoverlap_b200/lib/libfo_A.so timeout 60 python scripts/probe_ab.py blj256 2>&1 | tail -1 | cut -c1-220
timeout 60 python scripts/probe_ab.py blj256 2>&1 | tail -1 | cut -c1-220
