set -x
nvidia-smi -L
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/eval_format.py BC7 --size 1024 --big 8192 2>&1 | tail -5
python tools/eval_format.py BC4 --size 512 --big 8192 2>&1 | tail -5
python tools/eval_format.py BC5 --size 512 --big 8192 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_bc7.json
python bench.py --steps 5 --warmup 3 --format BC4 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_bc4.json
