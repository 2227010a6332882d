timeout 300 python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=1,2:pair_kernel=2:pair_segs=2,2:pair_kernel=2:pair_segs=4" 5 2>&1 | grep -E "mode|agrees|rror" | sed "s/^/main /"
for v in sleep64 sleep256 w4x6 w6x4; do
  FLASHFRY_B200_LIB=gpurun_variants/$v/libflashfry_b200.so timeout 300 python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=2:pair_segs=2,2:pair_kernel=2:pair_segs=4,2:pair_kernel=2:pair_segs=8" 5 2>&1 | grep -E "mode|rror" | sed "s/^/$v /"
done
