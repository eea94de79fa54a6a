#!/bin/bash
N=$1
OUT=gpurun_out/r02_n$N; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
wc -l $OUT/bench_n$N.json
python -c "import json;d=json.load(open('$OUT/bench_n$N.json'));print('n$N ms_per_step %.3f value %.3e e2e %.3e'%(d['ms_per_step'],d['value'],d['e2e']['value']))" || tail -5 $OUT/bench_n$N.err
