#!/bin/bash
# tools/ab_r2.sh tag -- A/B of the round-2 rollout changes on one box: emit position (early / late build), start stagger,
# host-call slices.  cfg2, K = 20 / 128 / 1000.
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # lib stagger K label
  FXENV_LIB=$1 FXENV_STAGGER_NS=$2 python bench.py --steps $3 --warmup 5 --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads 2>$OUT/err.txt | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-20s stagger=%-4s K=%-5d %7.2f us/step %7.1f M  e2e %.2f M' % ('$1', '$2', $3, d['ms_per_step']*1e3, d['value']/1e6, d['e2e']['value']/1e6))" || tail -3 $OUT/err.txt
}
for lib in libfxenv_early.so libfxenv_late.so; do for st in 0 300; do for k in 20 128 1000; do run $lib $st $k; done; done; done
for st in 150 600 1000; do for k in 20 128; do run libfxenv_early.so $st $k; done; done
for hs in 1 2 4; do echo "host slices $hs"; FXENV_HOST_SLICES=$hs run libfxenv_early.so 300 128; done
