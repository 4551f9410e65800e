out=gpurun_out; mkdir -p $out
( timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -x ) > $out/r02_mgpu_tests.log 2>&1; tail -3 $out/r02_mgpu_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > $out/r02_bench_c2_n2.json 2> $out/r02_bench_c2_n2.err; tail -c 400 $out/r02_bench_c2_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference ) > $out/r02_bench_c2_n2_ref.json 2> $out/r02_bench_c2_n2_ref.err
TRMF_B200_PACK_ORDER=split CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-parity --strong none --no-cpu-baseline > $out/r02_bench_pack_split.json 2> $out/r02_bench_pack_split.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-parity --strong none --no-cpu-baseline > $out/r02_bench_pack.json 2> $out/r02_bench_pack.err
