export LDAGPU_P2P_TIMEOUT_MS=10000
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python tests/singleproc_multigpu_check.py --gpus 2 --stress 40 2>&1 | tail -6 | cut -c1-700
