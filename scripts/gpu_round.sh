#!/bin/bash
# Evidence bundle for the CURRENT build, run on the GPU box:   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh r02_x'
#   gpurun_out/<tag>/pytest_gpu.log         the -m gpu suite
#   gpurun_out/<tag>/bench_n1.json, bench_reference_n1.json   both bench arms (CUDA-event timing, never under a profiler)
#   gpurun_out/<tag>/launches.csv           every launch of one SOL-32 iteration (ncu gpu__time_duration.sum)
#   gpurun_out/<tag>/ncu_<kernel>.json      `ncu --set full`, warm caches, per kernel: DRAM bytes, duration, tensor-pipe %, SASS histogram,
#                                           stamped with the source hash of this build (bench.py reads `traffic` from profiles/ncu_*.json)
# Copy what is to be judged into profiles/ (tracked); gpurun_out/ is scratch.
TAG=${1:-round}
WHAT=${2:-all}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi.txt 2>&1
if [[ $WHAT == all || $WHAT == *tests* ]]; then
  timeout 2400 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
fi
if [[ $WHAT == all || $WHAT == *bench* ]]; then
  python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err
  python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  python -c "import json;d=json.load(open('$OUT/bench_n1.json'));print('ms_per_step %.3f value %.3e e2e %.3e roofline frac %.4f (%s, %.2f us)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel'][:20],d['roofline']['us_per_launch']))"
fi
if [[ $WHAT == all || $WHAT == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv \
      python scripts/profile_iter.py --msteps 32 --graph > $OUT/launches.log 2>&1
  python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.md 2>/dev/null; head -30 $OUT/launches.md
  timeout 1500 ncu --set full --clock-control none --cache-control none --profile-from-start off -f -o $OUT/prof \
      python scripts/profile_iter.py --msteps 2 > $OUT/prof.log 2>&1
  python scripts/ncu_to_json.py $OUT/prof.ncu-rep $OUT "ncu --set full --clock-control none --cache-control none (warm caches), scripts/profile_iter.py --msteps 2, mean over the captured launches" > $OUT/ncu_kernels.md
  cat $OUT/ncu_kernels.md
  # gpurun merges at most 64 MiB back: keep the per-kernel JSON summaries, drop the report itself
  ls -la $OUT/prof.ncu-rep; rm -f $OUT/prof.ncu-rep
fi
