#!/bin/bash
OUT=gpurun_out/r02_n2b; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
head -c 120 $OUT/bench_n2.json; echo; wc -l $OUT/bench_n2.json
python -c "import json;d=json.load(open('$OUT/bench_n2.json'));print('n2 ms_per_step %.3f value %.3e e2e %.3e'%(d['ms_per_step'],d['value'],d['e2e']['value']))"
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2_info.json 2> $OUT/bench_n2_info.err
wc -l $OUT/bench_n2_info.json; grep -c NCCL $OUT/bench_n2_info.err; grep -i "NVLS\|via P2P" $OUT/bench_n2_info.err | head -3
