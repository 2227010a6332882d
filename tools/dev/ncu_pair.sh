ncu --set full --clock-control none --import-source on -k regex:k_pair_scan --launch-skip 2 --launch-count 1 -o gpurun_out/r3_pair_old -f python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=1" 4 > gpurun_out/ncu_pair_old.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pair_scan2 --launch-skip 2 --launch-count 1 -o gpurun_out/r3_pair_new -f python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=0" 4 > gpurun_out/ncu_pair_new.log 2>&1
tail -3 gpurun_out/ncu_pair_new.log
