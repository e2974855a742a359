set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12
python tools/eval_format.py ETC1 --size 512 --big 4096 2>&1 | tail -3
