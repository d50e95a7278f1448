# final per-round evidence run (one B200): bench line, ncu launch lists, ncu --set full captures
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err || exit 1
VLQ_PROFILE=search ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_search.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ls.log 2>&1
VLQ_PROFILE=encode ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_encode.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n 2000000 > gpurun_out/ncu_le.log 2>&1
VLQ_PROFILE=search ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"l2_tc|coarse_select|scan_topk|term3" -c 8 -o gpurun_out/r01_prof_search -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fs.log 2>&1
VLQ_PROFILE=encode ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"l2_tc|line_encode" -c 3 -o gpurun_out/r01_prof_encode -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --n 2000000 > gpurun_out/ncu_fe.log 2>&1
python bench.py --db-size 1000000000 --u8 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r01_bench_c4_1B_1gpu.json 2> gpurun_out/r01_bench_c4_1B_1gpu.err
tail -c 300 gpurun_out/r01_bench.json; tail -c 300 gpurun_out/r01_bench_c4_1B_1gpu.json
