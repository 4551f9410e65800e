out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:f_update_mma2 -c 1 -o $out/r02_mma2_k40_v2 -f tools/test_f_update_mma2 40 c2 > $out/r02_mma2_ncu.log 2>&1
tail -3 $out/r02_mma2_ncu.log
