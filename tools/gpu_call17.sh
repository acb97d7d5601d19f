#!/bin/bash
# driver-style end-of-round sequence on 2 GPUs: smoke, reference arm, default bench (N=1), torchrun N=2
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/c17_smoke.log 2>&1; tail -4 gpurun_out/c17_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/c17_bench_reference.json 2> gpurun_out/c17_bench_reference.err; tail -c 400 gpurun_out/c17_bench_reference.json; tail -3 gpurun_out/c17_bench_reference.err
( time python bench.py ) > gpurun_out/c17_bench_default.json 2> gpurun_out/c17_bench_default.err; tail -c 700 gpurun_out/c17_bench_default.json; tail -3 gpurun_out/c17_bench_default.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/c17_bench_n2.json 2> gpurun_out/c17_bench_n2.err; tail -c 900 gpurun_out/c17_bench_n2.json; tail -3 gpurun_out/c17_bench_n2.err
