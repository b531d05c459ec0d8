#!/bin/bash
# tools/mb_variants.sh lib1.so lib2.so ... -- bench cfg2 at 4096 and 16384 envs for each library variant (experiments)
for lib in "$@"; do for n in 4096 16384; do
  FXENV_LIB=$lib python bench.py --workload cfg2 --envs $n --steps 1000 --warmup 300 --no-cpu-baseline --no-single-step 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-22s envs %6d: %7.2f us/step  %7.1f M steps/s' % ('$lib', $n, d['ms_per_step']*1e3, d['value']/1e6))"
done; done
