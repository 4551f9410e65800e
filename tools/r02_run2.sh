set -u
out=gpurun_out; mkdir -p $out
( time timeout 900 python -m pytest tests/test_configs_gpu.py -m gpu -q -x -s --durations=10 ) > $out/r02_config_tests.log 2>&1; tail -5 $out/r02_config_tests.log
( time timeout 900 python bench.py ) > $out/r02_bench_c2_v1.json 2> $out/r02_bench_c2_v1.err; tail -c 600 $out/r02_bench_c2_v1.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $out/r02_bench_c2_ref_v1.json 2> $out/r02_bench_c2_ref_v1.err
free -g > $out/r02_mem.txt
echo done
