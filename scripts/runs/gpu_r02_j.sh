#!/bin/bash
mkdir -p gpurun_out/r02_j
O=gpurun_out/r02_j
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist_nccl.py -q -m gpu -s > $O/pytest_nccl.log 2>&1; tail -8 $O/pytest_nccl.log
timeout 900 python -m pytest tests/test_gpu_unroll.py tests/test_gpu_stages.py -q -m gpu -x > $O/pytest_sel.log 2>&1; tail -3 $O/pytest_sel.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; python -c "import json;d=json.load(open('$O/bench_n1.json'));print('n1 ms_per_step %.3f value %.4e'%(d['ms_per_step'],d['value']))"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --fuse-small 1 > $O/bench_n1_fs.json 2> $O/bench_n1_fs.err; python -c "import json;d=json.load(open('$O/bench_n1_fs.json'));print('n1 fuse_small ms_per_step %.3f'%(d['ms_per_step']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; python -c "import json;d=json.load(open('$O/bench_n2.json'));print('n2 ms_per_step %.3f value %.4e'%(d['ms_per_step'],d['value']))" || tail -5 $O/bench_n2.err
