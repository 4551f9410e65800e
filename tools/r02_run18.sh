out=gpurun_out; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > $out/r02_gpu_tests_mma2.log 2>&1; tail -25 $out/r02_gpu_tests_mma2.log
timeout 600 python bench.py --no-e2e > $out/r02_bench_c2_mma2_noe2e.json 2> $out/r02_bench_c2_mma2_noe2e.err; cut -c1-1500 $out/r02_bench_c2_mma2_noe2e.json; tail -3 $out/r02_bench_c2_mma2_noe2e.err
