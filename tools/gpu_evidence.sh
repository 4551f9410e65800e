#!/usr/bin/env bash
# gpu_evidence.sh -- everything the profiles/ directory is built from, in one call on a B200 box:
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh [quick|full] [round-tag]'
#
# quick (default, ~4 min): GPU test suite, smoke, bench lines for C2 / C3 (with parity, roofline_x, strong C5 record), reference arm,
#                          rolling_validate runs, ncu launch list of the bench command.
# ncu:                     only what `full` adds.
# full  (+ ~4 min):        also the ncu --set full captures (the kernels of one outer iteration of the default path: complement
#                          Gram, fp64 product, solve, Gram assembly; and the two Gram launches of the walk over Omega with
#                          TRMF_B200_COMPLEMENT=0), the standalone kernel test, the walk-path bench line and the C4 line.
# Outputs go to gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/ (names r<round>_*).
# Numbers printed by a command running under ncu are never bench values.
set -u
mode="${1:-quick}"
tag="${2:-r02}"
out=gpurun_out
mkdir -p "$out"
run() { local name="$1"; shift; echo "== $name: $*"; ( time timeout "${TMO:-300}" "$@" ) > "$out/$name.log" 2>&1; tail -4 "$out/$name.log"; }

if [ "$mode" != "ncu" ]; then
TMO=900 run ${tag}_gpu_tests python -m pytest tests -m gpu -q -x --durations=8
TMO=120 run ${tag}_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=120 run ${tag}_slab_sanity python tools/slab_sanity.py
timeout 400 python bench.py > "$out/${tag}_bench_c2_n1.json" 2> "$out/${tag}_bench_c2_n1.err"; cut -c1-300 "$out/${tag}_bench_c2_n1.json"
timeout 300 python bench.py --config c3 --strong none > "$out/${tag}_bench_c3_n1.json" 2> "$out/${tag}_bench_c3_n1.err"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > "$out/${tag}_bench_c2_reference_arm.json" 2> "$out/${tag}_bench_c2_reference_arm.err"
TRMF_B200_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-parity --strong none --no-cpu-baseline > /dev/null 2> "$out/${tag}_trace_c2_host_buffer_call.txt"
python tools/h2d_probe.py >> "$out/${tag}_trace_c2_host_buffer_call.txt" 2>&1
for cfg in electricity traffic; do
    timeout 200 python tools/rolling_bench.py --config "$cfg" --repeat 2 > "$out/${tag}_rolling_$cfg.json" 2> "$out/${tag}_rolling_$cfg.err"; cat "$out/${tag}_rolling_$cfg.json"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file "$out/${tag}_launches_c2.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-parity --strong none --no-cpu-baseline > "$out/${tag}_ncu_launches.log" 2>&1
python tools/ncu_summary.py "$out/${tag}_launches_c2.csv" > "$out/${tag}_launches_c2.txt"; head -14 "$out/${tag}_launches_c2.txt"
# the same with the host-buffer arm (c_trmf_train: per-slab placement / Gram / product / solve kernels of the first F-update)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$out/${tag}_launches_c2_e2e.csv" \
    python bench.py --steps 1 --warmup 1 --no-parity --strong none --no-cpu-baseline > "$out/${tag}_ncu_launches_e2e.log" 2>&1
python tools/ncu_summary.py "$out/${tag}_launches_c2_e2e.csv" > "$out/${tag}_launches_c2_e2e.txt"
fi

if [ "$mode" = "full" ] || [ "$mode" = "ncu" ]; then
    timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:gemm64|f_update_mma2|solve_kernel|xgram_kernel' -s 6 -c 6 \
        -o "$out/${tag}_complement_full" -f python bench.py --steps 1 --warmup 1 --no-e2e --no-parity --strong none --no-cpu-baseline > "$out/${tag}_ncu_full.log" 2>&1
    # (the reports are turned into text here: gpurun brings back at most 64 MiB)
    python tools/ncu_kernel_report.py "$out/${tag}_complement_full.ncu-rep" > "$out/${tag}_complement_ncu_full.txt" 2>&1; rm -f "$out/${tag}_complement_full.ncu-rep"
    TRMF_B200_COMPLEMENT=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:f_update_mma2 -s 2 -c 2 \
        -o "$out/${tag}_mma2_k40_full" -f python bench.py --steps 1 --warmup 1 --no-e2e --no-parity --strong none --no-cpu-baseline > "$out/${tag}_ncu_full2.log" 2>&1
    python tools/ncu_kernel_report.py "$out/${tag}_mma2_k40_full.ncu-rep" > "$out/${tag}_mma2_k40_ncu_full.txt" 2>&1; rm -f "$out/${tag}_mma2_k40_full.ncu-rep"
    TRMF_B200_COMPLEMENT=0 timeout 400 python bench.py --strong none --no-cpu-baseline > "$out/${tag}_bench_c2_n1_walk.json" 2> "$out/${tag}_bench_c2_n1_walk.err"
    for a in "40 small" "64 small" "20 small" "40 c2" "64 c5"; do echo "=== $a"; timeout 120 tools/test_f_update_mma2 $a 2>&1 | tail -8; done > "$out/${tag}_test_f_update_mma2.txt" 2>&1
    tools/microbench_dfma > "$out/${tag}_microbench_dfma.txt" 2>&1
    timeout 400 python bench.py --config c4 --no-e2e --strong none > "$out/${tag}_bench_c4_n1.json" 2> "$out/${tag}_bench_c4_n1.err"
fi
echo "== done ($mode)"
