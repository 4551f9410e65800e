out=gpurun_out; mkdir -p $out
for algo in words or8; do
echo "== algo $algo"
TRMF_B200_PACK_ALGO=$algo TRMF_B200_TRACE=1 TRMF_B200_TRACE_SYNC=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-parity > $out/r02_e2e_trace3.json 2> $out/r02_e2e_trace3.err
grep "trace" $out/r02_e2e_trace3.err | tail -10
done
for thr in 8 12 16; do
TRMF_B200_PACK_THREADS=$thr timeout 600 python bench.py --steps 9 --warmup 3 --no-parity > $out/r02_e2e_v3_$thr.json 2> $out/r02_e2e_v3.err
python -c "
import json; d=json.load(open('$out/r02_e2e_v3_$thr.json')); e=d['e2e']; print('threads $thr', e['ms_per_step'], e['ms_per_step_min'], e['ms_per_step_mean'])"
done
