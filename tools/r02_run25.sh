out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm64 -c 2 -o $out/r02_gemm64_v1 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-parity --strong none --no-cpu-baseline > $out/r02_gemm64_ncu.log 2>&1
tail -2 $out/r02_gemm64_ncu.log
