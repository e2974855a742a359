# Developer helper (GPU box, 1 GPU): the round's final measurements -> gpurun_out/r02f_* (copied into profiles/ by hand)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02f_pytest_gpu.txt
python bench.py --impl reference > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M -k regex:bc7_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02f_dram_BC7.csv python tools/prof_one.py BC7 8192 > /dev/null 2>&1
ncu --metrics $M -k regex:astc3 -s 2 -c 1 --csv --log-file gpurun_out/r02f_dram_ASTC_6x6.csv python tools/prof_one.py ASTC_6x6 8192 > /dev/null 2>&1
ncu --metrics $M -k regex:bc6h -s 2 -c 1 --csv --log-file gpurun_out/r02f_dram_BC6H.csv python tools/prof_one.py BC6H 4096 hdr > /dev/null 2>&1
ncu --metrics $M -k regex:etc_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02f_dram_ETC2_R8G8B8A8.csv python tools/prof_one.py ETC2_R8G8B8A8 4096 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:astc3 -s 1 -c 1 -o gpurun_out/r02f_astc6 -f python tools/prof_one.py ASTC_6x6 4096 > /dev/null 2>&1
python tools/real_report.py > gpurun_out/r02f_real_report.txt 2>&1
tail -3 gpurun_out/r02f_pytest_gpu.txt; cat gpurun_out/r02f_bench_n1.json | cut -c1-600
