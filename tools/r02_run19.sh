out=gpurun_out; mkdir -p $out
TRMF_B200_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-parity > $out/r02_e2e_trace.json 2> $out/r02_e2e_trace.err
grep "trace" $out/r02_e2e_trace.err | tail -28
TRMF_B200_TRACE=1 TRMF_B200_TRACE_SYNC=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-parity > $out/r02_e2e_trace2.json 2> $out/r02_e2e_trace2.err
grep "trace" $out/r02_e2e_trace2.err | tail -15
python -c "
import json; d=json.load(open('$out/r02_e2e_trace.json')); print(d['e2e'])"
