#!/bin/bash
OUT=gpurun_out/r02_n2; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -q -m gpu -s > $OUT/pytest_nccl.log 2>&1; tail -6 $OUT/pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
python -c "import json;d=json.load(open('$OUT/bench_n2.json'));print('n2 ms_per_step %.3f value %.3e e2e %.3e'%(d['ms_per_step'],d['value'],d['e2e']['value']))"
