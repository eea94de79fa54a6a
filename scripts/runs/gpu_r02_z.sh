#!/bin/bash
OUT=gpurun_out/r02_z; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_quoted_configs.py tests/test_gpu_compat.py -x -q -m gpu > $OUT/pytest_sel.log 2>&1; tail -5 $OUT/pytest_sel.log
grep -h "weight gradient rel" gpurun_out/parity_*.json 2>/dev/null | head -3
python - <<'P'
import json
for n in ("sol32","c2","c4"):
    try:
        d=json.load(open("gpurun_out/parity_%s.json"%n)); print(n,"grad",d["grad"],"layers dW",[round(l["dW"]*1e5,2) for l in d["grad_layers"]])
    except Exception as e: print(n,e)
P
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; head -6 $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
timeout 300 python scripts/chain_trace.py --opt wgrad_pair=0 > $OUT/chain_trace_sol32_nopair.txt 2>> $OUT/chain_trace.err; head -6 $OUT/chain_trace_sol32_nopair.txt | tail -3
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "import json;d=json.load(open('$OUT/bench.json'));print('ms_per_step %.3f e2e %.3f conv %.2f us'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['us_per_launch']))"
