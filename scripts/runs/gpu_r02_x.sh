#!/bin/bash
OUT=gpurun_out/r02_x; mkdir -p $OUT
for v in 3 4 0; do
timeout 300 python scripts/chain_trace.py --opt conv_variant=$v > $OUT/chain_trace_sol32_v$v.txt 2> $OUT/chain_trace.err; head -5 $OUT/chain_trace_sol32_v$v.txt | tail -3; tail -3 $OUT/chain_trace.err
done
