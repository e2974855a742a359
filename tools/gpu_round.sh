set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
python tools/eval_format.py ASTC_6x6 --size 516 --big 4096 2>&1 | tail -3
python tools/eval_format.py BC7 --size 1024 --big 8192 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:astc -s 1 -c 1 -f -o gpurun_out/prof_astc python tools/prof_one.py ASTC_6x6 1536 noise+grad 2 > gpurun_out/ncu_astc.log 2>&1
tail -2 gpurun_out/ncu_astc.log
