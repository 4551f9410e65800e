out=gpurun_out; mkdir -p $out
( time timeout 900 python -m pytest tests/test_rolling_gpu.py tests/test_ingest_gpu.py -m gpu -q --durations=5 ) > $out/r02_roll_ingest_tests.log 2>&1; tail -12 $out/r02_roll_ingest_tests.log
for cfg in electricity traffic; do
    timeout 200 python tools/rolling_bench.py --config "$cfg" --repeat 2 > "$out/r02_rolling_$cfg.json" 2> "$out/r02_rolling_$cfg.err"; cut -c1-700 "$out/r02_rolling_$cfg.json"
done
