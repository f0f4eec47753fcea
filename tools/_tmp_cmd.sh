python tools/diag_chain.py 100
python tools/diag_host.py 2>&1 | tail -7
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
