set -x
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err || exit 1
python bench.py --impl reference > gpurun_out/r01_bench_ref.json 2> gpurun_out/r01_bench_ref.err
VLQ_PROFILE=search ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_search.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ls.log 2>&1
VLQ_PROFILE=encode ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_encode.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n 2000000 > gpurun_out/ncu_le.log 2>&1
VLQ_PROFILE=search ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"l2_tc|coarse_select|scan_topk|term3" -c 8 -o gpurun_out/r01_prof_search -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fs.log 2>&1
VLQ_PROFILE=encode ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"l2_tc|line_encode" -c 3 -o gpurun_out/r01_prof_encode -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n 2000000 > gpurun_out/ncu_fe.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scan_async -s 4 -c 1 -o gpurun_out/r01_prof_scan_1B -f python bench.py --db-size 1000000000 --u8 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_1b.log 2>&1
tail -c 600 gpurun_out/r01_bench.json; tail -c 400 gpurun_out/r01_bench_ref.json
