out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --no-parity --strong none --no-cpu-baseline > $out/r02_bench_pack.json 2> $out/r02_bench_pack.err
TRMF_B200_PACK_ORDER=split timeout 600 python bench.py --no-parity --strong none --no-cpu-baseline > $out/r02_bench_pack_split.json 2> $out/r02_bench_pack_split.err
( timeout 600 python -m pytest tests/test_ingest_gpu.py -m gpu -q -x ) > $out/r02_ingest_tests.log 2>&1; tail -2 $out/r02_ingest_tests.log
