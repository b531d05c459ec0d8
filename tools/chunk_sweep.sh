#!/bin/bash
# tools/chunk_sweep.sh -- cfg2 step_many time vs steps-per-ticket (FXENV_CHUNK) at several batch lengths
# usage: chunk_sweep.sh tag "K:c1,c2,..." ...
OUT=gpurun_out/${1:-chunk}; mkdir -p $OUT; shift
for spec in "$@"; do k=${spec%%:*}; for ch in $(echo ${spec#*:} | tr , ' '); do
  FXENV_CHUNK=$ch timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads > $OUT/k${k}_c$ch.json 2> $OUT/k${k}_c$ch.err
  python -c "import json; d=json.load(open('$OUT/k${k}_c$ch.json')); print('K=%-5d chunk=%-3d %7.2f us/step  frac %.3f' % ($k, $ch, d['ms_per_step']*1e3, d['roofline']['frac']))"
done; done
