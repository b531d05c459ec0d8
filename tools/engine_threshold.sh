#!/bin/bash
# tools/engine_threshold.sh -- persistent launch (FXENV_DEBUG=16) vs graph of single steps (8) for cfg2-shaped workloads of
# growing size: where does the library's size rule (rollout while envs <= 3 x resident warps) stop being right?
for n in 2048 4096 6144 8192 12288 16384; do for mode in 16 8; do
  FXENV_DEBUG=$mode python bench.py --workload cfg2 --envs $n --steps 1000 --warmup 300 --no-cpu-baseline --no-single-step 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('envs %6d mode=%-2s %7.2f us/step  %7.1f M steps/s' % ($n, '$mode', d['ms_per_step']*1e3, d['value']/1e6))"
done; done
