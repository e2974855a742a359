set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15
for f in ETC1 ETC2_R8G8B8A8; do python tools/eval_format.py $f --size 512 --big 4096 2>&1 | tail -3; done
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bc7.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bc7.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
