set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/eval_format.py BC7 --size 1024 --big 8192 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:bc7 -s 1 -c 1 -f -o gpurun_out/prof_bc7 python tools/prof_one.py BC7 4096 > gpurun_out/ncu_bc7.log 2>&1
tail -3 gpurun_out/ncu_bc7.log
