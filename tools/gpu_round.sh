set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
for f in BC1_RGB BC3; do python tools/eval_format.py $f --size 1024 --big 8192 2>&1 | tail -3; done
python tools/eval_format.py BC6H --type UFloat --kinds hdr --size 512 --big 4096 2>&1 | tail -2
