# final per-round evidence run (one B200): smoke, bench line, ncu launch lists, ncu --set full captures
# usage (under gpurun): bash tools/final_profile.sh r02
R=${1:-r02}
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err || exit 1
VLQ_PROFILE=search ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_search.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4-stage > gpurun_out/ncu_ls.log 2>&1
VLQ_PROFILE=encode ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_encode.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c4-stage --n 2000000 > gpurun_out/ncu_le.log 2>&1
VLQ_PROFILE=search ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"l2_tc|coarse_select|scan_topk|term3" -c 8 -o gpurun_out/${R}_prof_search -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c4-stage > gpurun_out/ncu_fs.log 2>&1
# the scan kernel at BASELINE configs[3] list density (synthetic lists, tools/bench_scan.py)
python tools/bench_scan.py --nq 2048 --steps 1 > gpurun_out/${R}_scan_c4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_long -s 1 -c 1 -o gpurun_out/${R}_prof_scan_long -f python tools/bench_scan.py --nq 2048 --steps 1 > gpurun_out/ncu_sl.log 2>&1
python tools/bench_scan.py --steps 5 > gpurun_out/${R}_scan_c4.json 2>&1
tail -c 400 gpurun_out/${R}_bench.json; tail -c 400 gpurun_out/${R}_scan_c4.json
