out=gpurun_out; mkdir -p $out; rm -f $out/r02_tc_exp.txt
for cfg in "64 112" "64 48" "128 16" "128 112" "64 16"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTRMF_F32=1 -DValueType=float -DTC_EXP_M=$1 -DTC_EXP_N=$2 -o tools/test_f_update_tc_exp tools/test_f_update_tc.cu 2>> $out/r02_tc_build.log
  echo "== M=$1 N=$2" >> $out/r02_tc_exp.txt
  TC_CLK=1 timeout 60 tools/test_f_update_tc_exp 40 c2 2>&1 | grep -E "warp  8:|MODE_DEFER  tcgen05" | head -2 >> $out/r02_tc_exp.txt
done
