out=gpurun_out; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTRMF_F32=1 -DValueType=float -o tools/test_f_update_tc tools/test_f_update_tc.cu 2> $out/r02_tc_build.log
timeout 60 tools/test_f_update_tc 40 c2 > $out/r02_test_f_update_tc.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:f_update_tc_kernel -s 3 -c 1 -o $out/r02_tc_k40_v2 -f tools/test_f_update_tc 40 c2 > $out/r02_tc_ncu.log 2>&1
tail -3 $out/r02_tc_ncu.log
