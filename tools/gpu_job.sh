#!/bin/bash
# One parametrised GPU-box job (replaces the per-round gpu_*.sh scripts).  Usage, from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_job.sh tests configs bench'
# Stages (run in the order given; every stage writes under gpurun_out/):
#   tests        pytest -m gpu
#   bench        bench.py (N = 1) and the reference arm
#   configs      tools/bench_configs.py: every BASELINE.json config, device-resident
#   launches     ncu launch list of a short bench.py run (never a bench value)
#   ncu:<tag>:<kernel regex>:<ncu_target.py args with , for spaces>   one `ncu --set full` capture
#   variants     tools/variants.py run: the -D ablation libraries built beforehand with `tools/variants.py build`
#   gate2ks      tools/validate_gate2.py ks     gate2ab:<dim>:<T>:<n>   tools/validate_gate2.py ab
#   py:<script>[:args,...]   any tools/*.py script, output to gpurun_out/<script>.txt
mkdir -p gpurun_out
for stage in "$@"; do
  echo "=== stage $stage ($(date +%T))"
  case "$stage" in
    tests)    timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 25 | tee gpurun_out/pytest_gpu.txt ;;
    bench)    timeout 900 python bench.py 2>&1 | tee gpurun_out/bench.txt
              timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref.txt ;;
    configs)  timeout 900 python tools/bench_configs.py 2>&1 | tee gpurun_out/bench_all_configs.jsonl ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
                python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dat > gpurun_out/bench_under_ncu.log 2>&1
              tail -n 3 gpurun_out/bench_under_ncu.log ;;
    ncu:*)    IFS=: read -r _ tag regex targs <<< "$stage"
              timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s 1 -c 1 -f \
                -o "gpurun_out/prof_$tag" python tools/ncu_target.py ${targs//,/ } > "gpurun_out/ncu_$tag.log" 2>&1
              tail -n 2 "gpurun_out/ncu_$tag.log"
              # reports are ~17 MB each and gpurun_out/ is capped at 64 MiB: digest on the box, keep the report only on request
              python tools/ncu_digest.py "gpurun_out/prof_$tag.ncu-rep" > "gpurun_out/ncu_digest_$tag.txt" 2>&1
              if [ -z "$KEEP_REP" ]; then rm -f "gpurun_out/prof_$tag.ncu-rep"; fi ;;
    variants) timeout 1200 python tools/variants.py run 2>&1 | tee gpurun_out/variants.txt ;;
    gate2ks)  timeout 1200 python tools/validate_gate2.py ks 2>&1 | tee gpurun_out/gate2_ks.txt ;;
    gate2ab:*) IFS=: read -r _ dim T n <<< "$stage"
              timeout 1500 python tools/validate_gate2.py ab --dim "$dim" --T "$T" --n "$n" 2>&1 | tee "gpurun_out/gate2_ab_dim$dim.txt" ;;
    py:*)     IFS=: read -r _ script sargs <<< "$stage"
              timeout 1500 python "tools/$script.py" ${sargs//,/ } 2>&1 | tee "gpurun_out/$script.txt" ;;
    *)        echo "unknown stage $stage" ;;
  esac
done
ls -la gpurun_out | tail -n 30
