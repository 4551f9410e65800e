out=gpurun_out; mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q --durations=10 ) > $out/r02_gpu_tests.log 2>&1; tail -25 $out/r02_gpu_tests.log
