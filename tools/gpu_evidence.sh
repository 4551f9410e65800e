#!/usr/bin/env bash
# gpu_evidence.sh -- everything the profiles/ directory is built from, in one call on a B200 box:
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_evidence.sh [quick|full]'
#
# quick (default, ~3 min): GPU test suite, smoke, bench lines for C2 / C3, rolling_validate runs, per-phase probe,
#                          ncu launch list of the bench command.
# full  (+ ~4 min):        also the CG-control microbenchmark, the ncu --set full capture of the Gram kernel and the
#                          C4 bench line.
# Outputs go to gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/ (names r<round>_*).
# Numbers printed by a command running under ncu are never bench values.
set -u
mode="${1:-quick}"
out=gpurun_out
mkdir -p "$out"
run() { local name="$1"; shift; echo "== $name: $*"; ( time timeout "${TMO:-300}" "$@" ) > "$out/$name.log" 2>&1; tail -4 "$out/$name.log"; }

TMO=300 run gpu_tests python -m pytest tests -m gpu -q -x --durations=5
TMO=120 run smoke python -c "import __graft_entry__ as g; g.smoke()"
timeout 200 python bench.py > "$out/bench_c2_n1.json" 2> "$out/bench_c2_n1.err"; cut -c1-400 "$out/bench_c2_n1.json"
timeout 200 python bench.py --config c3 > "$out/bench_c3_n1.json" 2> "$out/bench_c3_n1.err"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_c2_reference_arm.json" 2> "$out/bench_c2_reference_arm.err"
for cfg in electricity traffic; do
    timeout 200 python tools/rolling_bench.py --config "$cfg" --repeat 2 > "$out/rolling_$cfg.json" 2> "$out/rolling_$cfg.err"; cat "$out/rolling_$cfg.json"
done
timeout 200 python tools/rolling_bench.py --config c2 --windows 3 --max-iter 10 --repeat 1 > "$out/rolling_c2.json" 2> "$out/rolling_c2.err"
timeout 100 python tools/rolling_phase_probe.py electricity traffic > "$out/rolling_phases.json" 2> "$out/rolling_phases.err"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file "$out/launches_c2.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e > "$out/ncu_launches.log" 2>&1
python tools/ncu_summary.py "$out/launches_c2.csv" > "$out/launches_c2.txt"; head -12 "$out/launches_c2.txt"

if [ "$mode" = "full" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_tcgen05_gram tools/microbench_tcgen05_gram.cu && \
        timeout 120 tools/microbench_tcgen05_gram > "$out/microbench_tcgen05_gram.txt" 2>&1; cat "$out/microbench_tcgen05_gram.txt"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/proto_f_update_tc tools/proto_f_update_tc.cu && \
        timeout 120 tools/proto_f_update_tc > "$out/proto_f_update_tc.txt" 2>&1; cat "$out/proto_f_update_tc.txt"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_cg tools/microbench_cg.cu && \
        timeout 120 tools/microbench_cg > "$out/microbench_cg.txt" 2>&1; cat "$out/microbench_cg.txt"
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:f_update_mma -c 4 -o "$out/f_update_mma_full" -f \
        python bench.py --steps 1 --warmup 1 --no-e2e > "$out/ncu_full.log" 2>&1
    timeout 300 python bench.py --config c4 --no-e2e > "$out/bench_c4_n1.json" 2> "$out/bench_c4_n1.err"
fi
echo "== done ($mode)"
