out=gpurun_out; mkdir -p $out
for v in 0 1 2 3 4 5; do
  echo "=== variant $v (0: 4 CTA x 2 stages, 1: 3x3, 2: 2x5, 3: 3x2, 4: 2x3, 5: 8 warps x 1 CTA x 5 stages)"; FM2_VARIANT=$v timeout 120 tools/test_f_update_mma2 40 c2 2>&1 | grep "mma2"
done > $out/r02_test_f_update_mma2_v2.txt 2>&1
cat $out/r02_test_f_update_mma2_v2.txt
