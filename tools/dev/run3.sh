timeout 300 python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=1,2:pair_kernel=2" 5 2>&1 | grep -E "mode|agrees|rror" | sed "s/^/main /"
for v in imad80 imad96 lop96; do
  FLASHFRY_B200_LIB=gpurun_variants/$v/libflashfry_b200.so timeout 300 python tools/scan_probe.py 3e8 100000 4 "2:pair_kernel=1,2:pair_kernel=2" 5 2>&1 | grep -E "mode|agrees|rror" | sed "s/^/$v /"
done
