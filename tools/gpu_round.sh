set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12
for f in BC1_RGB BC3; do python tools/eval_format.py $f --size 1024 --big 8192 2>&1 | tail -3; done
