#!/bin/bash
# timing decomposition of the step kernel (debug switches); prints us/step for cfg2 at several env counts
OUT=gpurun_out/${1:-mb}; mkdir -p $OUT
run() { # label, env assignments...
  label=$1; shift
  for n in 4096 16384; do
    env "$@" timeout 120 python bench.py --workload cfg2 --envs $n --steps 1000 --warmup 300 --no-cpu-baseline 2>>$OUT/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s envs %6d: %7.2f us/step  %7.1f M steps/s  frac %.3f' % ('$label', $n, d['ms_per_step']*1e3, d['value']/1e6, d['roofline']['frac']))"
  done
}
run full            FXENV_DEBUG=0
run no_obs          FXENV_DEBUG=1
run obs_only        FXENV_DEBUG=2
run cursor_only     FXENV_DEBUG=3
