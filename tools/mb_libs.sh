#!/bin/bash
# tools/mb_libs.sh lib1.so ... -- A/B of library builds: cfg2 (persistent + single-step graph), cfg3, cfg4
for lib in "$@"; do
  for spec in "cfg2 0" "cfg2 8" "cfg3 0" "cfg4 0"; do set -- $spec
    FXENV_LIB=$lib FXENV_DEBUG=$2 python bench.py --workload $1 --steps 1000 --warmup 300 --no-cpu-baseline --no-single-step 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-18s %-5s debug=%-2s %7.2f us/step  %7.1f M steps/s' % ('$lib', '$1', '$2', d['ms_per_step']*1e3, d['value']/1e6))"
  done
done
