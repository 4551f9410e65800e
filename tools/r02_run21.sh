out=gpurun_out; mkdir -p $out
( nproc; grep -m1 "model name" /proc/cpuinfo; python tools/packbench.py 10000; g++ -O3 -o /tmp/pb tools/packbench_standalone.cpp && /tmp/pb ) > $out/r02_packbench.txt 2>&1
cat $out/r02_packbench.txt
