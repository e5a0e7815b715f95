set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python scripts/exp_cfg4.py
