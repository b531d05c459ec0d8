#!/bin/bash
# tools/mb_modes.sh -- bench every workload in the three step_many modes (rollout kernel / chained graph / grid-serialised graph)
for wl in cfg2 cfg3 cfg4 cfg5; do for mode in 0 16 8; do
  FXENV_DEBUG=$mode python bench.py --workload $wl --steps 1000 --warmup 300 --no-cpu-baseline --no-single-step 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-5s debug=%-2s envs %6d: %7.2f us/step  %7.1f M steps/s' % ('$wl', '$mode', d['config']['envs_per_gpu'], d['ms_per_step']*1e3, d['value']/1e6))"
done; done
