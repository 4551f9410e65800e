out=gpurun_out; mkdir -p $out
( for v in 0 1 2; do
  echo "=== .ca variant $v"; FM2_VARIANT=$v timeout 120 tools/test_f_update_mma2_ca 40 c2 2>&1 | grep "mma"
done
echo "=== .ca k=64 c5";  timeout 120 tools/test_f_update_mma2_ca 64 c5 2>&1 | grep "mma"
echo "=== .ca k=40 small";  timeout 120 tools/test_f_update_mma2_ca 40 small 2>&1 | grep "mma" ) > $out/r02_test_f_update_mma2_v3.txt 2>&1
cat $out/r02_test_f_update_mma2_v3.txt
