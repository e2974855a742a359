set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for q in Lowest Normal High Highest; do python tools/eval_format.py BC7 --size 1024 --big 8192 --quality $q 2>&1 | tail -3; done
ncu --set full --clock-control none --import-source on -k regex:bc7 -s 1 -c 1 -f -o gpurun_out/prof_bc7 python tools/prof_one.py BC7 4096 > gpurun_out/ncu_bc7.log 2>&1
tail -2 gpurun_out/ncu_bc7.log
