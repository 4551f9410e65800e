out=gpurun_out; mkdir -p $out
( time timeout 900 python -m pytest tests/test_complement_gpu.py tests/test_rolling_gpu.py tests/test_ingest_gpu.py -m gpu -q -x ) > $out/r02_tests_c.log 2>&1; tail -6 $out/r02_tests_c.log | cut -c1-200
timeout 600 python bench.py --strong none > $out/r02_bench_c2_compl2.json 2> $out/r02_bench_c2_compl2.err; tail -3 $out/r02_bench_c2_compl2.err
python -c "
import json; d=json.load(open('$out/r02_bench_c2_compl2.json')); print(d['ms_per_step'], d['phase_ms'], d['parity']['H'], d['cg_steps'][:3]); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline'])"
timeout 600 python bench.py --config c3 --strong none --no-cpu-baseline > $out/r02_bench_c3_compl2.json 2> $out/r02_bench_c3_compl2.err
python -c "
import json; d=json.load(open('$out/r02_bench_c3_compl2.json')); print('c3', d['ms_per_step'], d['phase_ms'], d['parity']['H'], d['parity']['W'], d['cg_steps'][:3]); print(d['e2e']['ms_per_step'])"
for cfg in traffic; do
    timeout 200 python tools/rolling_bench.py --config "$cfg" --repeat 2 > "$out/r02_rolling_${cfg}_compl.json" 2> "$out/r02_rolling_${cfg}_compl.err"; cut -c1-900 "$out/r02_rolling_${cfg}_compl.json"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_c2_compl.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-parity --strong none --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py $out/r02_launches_c2_compl.csv > $out/r02_launches_c2_compl.txt 2>&1; head -12 $out/r02_launches_c2_compl.txt
