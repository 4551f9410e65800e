out=gpurun_out; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > $out/r02_gpu_tests_compl.log 2>&1; tail -30 $out/r02_gpu_tests_compl.log | cut -c1-250
timeout 600 python bench.py --no-e2e > $out/r02_bench_c2_compl.json 2> $out/r02_bench_c2_compl.err; tail -3 $out/r02_bench_c2_compl.err
python -c "
import json; d=json.load(open('$out/r02_bench_c2_compl.json')); print(d['ms_per_step'], d['phase_ms'], d['parity'], d['cg_steps'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_c2_compl.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-parity --strong none --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py $out/r02_launches_c2_compl.csv > $out/r02_launches_c2_compl.txt 2>&1; head -24 $out/r02_launches_c2_compl.txt
