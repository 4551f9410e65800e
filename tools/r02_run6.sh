out=gpurun_out; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTRMF_F32=1 -DValueType=float -o tools/test_f_update_tc tools/test_f_update_tc.cu 2> $out/r02_tc_build.log
TC_CLK=1 timeout 60 tools/test_f_update_tc 40 c2 > $out/r02_test_f_update_tc_clk.txt 2>&1; echo "exit $?" >> $out/r02_test_f_update_tc_clk.txt
