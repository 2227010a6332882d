set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 300 python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=1,2:pair_kernel=0,2:pair_segs=1,2:pair_segs=2,2:pair_segs=8" 5 2>&1 | grep -E "mode|agrees|Error|error" 
for v in p2_2x320 p2_4x192; do
  FLASHFRY_B200_LIB=gpurun_variants/$v/libflashfry_b200.so timeout 300 python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=1,2:pair_kernel=0,2:pair_segs=2,2:pair_segs=8" 5 2>&1 | grep -E "mode|agrees|rror" | sed "s/^/$v /"
done
timeout 300 python tools/scan_probe.py 3e8 12500 4 "2:pair_kernel=1,2:pair_kernel=0,1" 5 2>&1 | grep -E "mode|agrees|rror"
timeout 300 python tools/scan_probe.py 3e8 40000 4 "2:pair_kernel=1,2:pair_kernel=0" 5 2>&1 | grep -E "mode|agrees|rror"
