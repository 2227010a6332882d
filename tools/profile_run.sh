# The profile capture of the round (one B200): gpurun -- bash tools/profile_run.sh ; summaries: tools/make_scan_traffic.py, tools/ncu_summary.py
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; tail -3 gpurun_out/r3_pytest.log
python bench.py --impl reference > gpurun_out/r3_ref.json 2> gpurun_out/r3_ref.err
python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -c 300 gpurun_out/r3_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r3_launches_bench.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_bin_scan|k_pair_scan" --launch-skip 4 --launch-count 2 -o gpurun_out/r3_scan -f python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2> gpurun_out/r3_ncu.log
ncu --set full --clock-control none --import-source on -k regex:"k_pair_scan2" --launch-skip 2 --launch-count 1 -o gpurun_out/r3_ring -f python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=2" 4 > gpurun_out/r3_ring.log 2>&1
ncu --set full --clock-control none -k regex:"k_sort_cut|k_guide_place|k_compact_rows" --launch-skip 6 --launch-count 3 -o gpurun_out/r3_order -f python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>> gpurun_out/r3_ncu.log
(compute-sanitizer --tool memcheck python tools/scan_probe.py 3e6 2000 4 "2:pair_kernel=1,2:pair_kernel=2" 1 2>&1 | grep -E "mode|ERROR SUMMARY|COMPUTE-SANITIZER" ; compute-sanitizer --tool racecheck python tools/scan_probe.py 3e6 300 4 "2:pair_kernel=2" 1 2>&1 | grep -E "mode|RACECHECK SUMMARY|COMPUTE-SANITIZER"; compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_multi.py -q -k "devices0" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|COMPUTE-SANITIZER") > gpurun_out/r3_sanitizer.txt 2>&1
cat gpurun_out/r3_sanitizer.txt
