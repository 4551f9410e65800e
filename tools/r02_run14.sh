out=gpurun_out; mkdir -p $out
for a in "40 small" "64 small" "20 small" "24 small" "48 small" "56 small" "60 small" "8 small" "40 c2" "64 c5"; do
  echo "=== $a"; timeout 120 tools/test_f_update_mma2 $a 2>&1 | tail -12
done > $out/r02_test_f_update_mma2_v5.txt 2>&1
grep -A5 "=== 40 small\|c2\|c5" $out/r02_test_f_update_mma2_v5.txt; grep -c BAD $out/r02_test_f_update_mma2_v5.txt
