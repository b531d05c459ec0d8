#!/bin/bash
# tools/gpu_check.sh -- one gpurun call: smoke + GPU tests + bench + ncu launch list + one full ncu capture.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag] [what...]'   what: smoke tests bench ncu
set -u
TAG=${1:-r1}; shift || true
WHAT=${*:-smoke tests bench ncu}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
for w in $WHAT; do case $w in
smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" ;;
tests) timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest.log ;;
ksweep) for k in 20 32 128 1000; do
         timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads > $OUT/bench_cfg2_k$k.json 2> $OUT/bench_cfg2_k$k.err
         python -c "import json,sys; d=json.load(open('$OUT/bench_cfg2_k$k.json')); print('K=%-5d %7.2f us/step  %7.1f M steps/s  frac %.3f  e2e %.2f M' % ($k, d['ms_per_step']*1e3, d['value']/1e6, d['roofline']['frac'], d['e2e']['value']/1e6))"; done ;;
bench) timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_cfg2_driver.json 2> $OUT/bench_cfg2_driver.err; echo "bench driver-like rc=$?"; tail -c 900 $OUT/bench_cfg2_driver.json
       for wl in cfg2 cfg3 cfg4 cfg5; do
         timeout 300 python bench.py --workload $wl --steps 1000 --warmup 100 --no-closed-loop --no-other-workloads > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl rc=$?"; tail -c 1500 $OUT/bench_$wl.json; done
       timeout 300 python bench.py --impl reference --steps 200 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json ;;
ncu)   timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
         python bench.py --steps 1000 --warmup 500 --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
       timeout 600 ncu --set full --clock-control none --import-source on -k regex:fx_rollout -s 1 -c 1 -f -o $OUT/prof_rollout \
         python bench.py --steps 200 --warmup 100 --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
       FXENV_DEBUG=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fx_step -s 340 -c 1 -f -o $OUT/prof_step \
         python bench.py --steps 100 --warmup 300 --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads > $OUT/ncu_full2.log 2>&1; echo "ncu full (single step) rc=$?" ;;
esac; done
ls -la $OUT
