set -x
python -m pytest tests -m gpu -q -k astc 2>&1 | tail -8
python tools/eval_format.py ASTC_6x6 --size 516 --big 4096 2>&1 | tail -3
CFX_ASTC_V1=1 python tools/eval_format.py ASTC_6x6 --size 516 --big 1024 --no-oracle 2>&1 | tail -3
compute-sanitizer --tool memcheck python tools/prof_one.py ASTC_6x6 192 noise+grad 1 2>&1 | tail -5
