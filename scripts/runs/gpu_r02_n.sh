#!/bin/bash
OUT=gpurun_out/r02_n; mkdir -p $OUT
timeout 300 python scripts/chain_trace.py > $OUT/chain_trace_sol32.txt 2> $OUT/chain_trace.err; cat $OUT/chain_trace_sol32.txt; tail -3 $OUT/chain_trace.err
