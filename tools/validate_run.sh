set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; tail -3 gpurun_out/r3_pytest.log
python bench.py --impl reference > gpurun_out/r3_ref.json 2> gpurun_out/r3_ref.err
python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -c 300 gpurun_out/r3_bench.err
python tools/scan_probe.py 3e8 200000 4 "2:pair_kernel=1,2:pair_kernel=2" 5 2>&1 | grep -E "mode|agrees" | sed -E 's/hits.*//'
python tools/scan_probe.py 3e8 400000 4 "2:pair_kernel=1,2:pair_kernel=2" 3 2>&1 | grep -E "mode|agrees" | sed -E 's/hits.*//'
python __graft_entry__.py smoke 2>&1 | tail -2
