out=gpurun_out; mkdir -p $out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "variants_agree or every_rank" ) > $out/r02_tc_parity_tests.log 2>&1; tail -4 $out/r02_tc_parity_tests.log
TRMF_B200_F_KERNEL=tc timeout 600 python bench.py --no-cpu-baseline --strong none --no-e2e > $out/r02_bench_c2_tc.json 2> $out/r02_bench_c2_tc.err; tail -c 300 $out/r02_bench_c2_tc.err
TRMF_B200_F_KERNEL=tc timeout 600 python bench.py --config c5 --no-cpu-baseline --strong none --no-e2e --steps 3 --warmup 2 > $out/r02_bench_c5_tc.json 2> $out/r02_bench_c5_tc.err; tail -c 300 $out/r02_bench_c5_tc.err
timeout 600 python bench.py --config c5 --no-cpu-baseline --strong none --no-e2e --steps 3 --warmup 2 > $out/r02_bench_c5_mma.json 2> $out/r02_bench_c5_mma.err
