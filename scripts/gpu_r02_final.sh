#!/bin/bash
# Final evidence bundle of round 2 (one B200): the gpu_round.sh bundle + chain trace + C2 / C4 bench lines + memcheck of the kernels added this round
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/gpu_round.sh $TAG all
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; head -24 $OUT/chain_trace_sol32.txt
python bench.py --config c2 --steps 20 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err
python bench.py --config c4 --steps 10 --warmup 3 > $OUT/bench_c4.json 2> $OUT/bench_c4.err
python -c "
import json
for n in ('c2','c4'):
    d=json.load(open('$OUT/bench_%s.json'%n)); print(n,'ms_per_step %.3f value %.3e'%(d['ms_per_step'],d['value']))"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_stages.py tests/test_gpu_unroll.py -m gpu -q -x \
    -k "any_grid or thin_paths or (64x32m2 and defaults and direct) or recaptured" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -3 $OUT/memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|Error" $OUT/memcheck.log | head -10
