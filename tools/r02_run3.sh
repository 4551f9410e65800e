out=gpurun_out; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench_gather tools/microbench_gather.cu
timeout 90 tools/microbench_gather > $out/r02_microbench_gather.txt 2>&1; echo "exit $?" >> $out/r02_microbench_gather.txt
timeout 90 tools/microbench_gather verify box4 > $out/r02_microbench_gather_box4.txt 2>&1; echo "exit $?" >> $out/r02_microbench_gather_box4.txt
nvidia-smi --query-gpu=name,clocks.sm --format=csv >> $out/r02_microbench_gather.txt
