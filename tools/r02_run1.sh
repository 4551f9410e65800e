set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/r02_box.txt; nproc >> $out/r02_box.txt; lscpu | head -20 >> $out/r02_box.txt
for t in microbench_tcgen05_gram proto_f_update_tc microbench_cg; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/$t tools/$t.cu 2> $out/r02_$t.build.log
  timeout 120 tools/$t > $out/r02_$t.txt 2>&1; echo "exit $?" >> $out/r02_$t.txt
done
timeout 300 python bench.py > $out/r02_bench_c2_start.json 2> $out/r02_bench_c2_start.err
echo done
