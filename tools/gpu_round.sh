#!/bin/bash
# One GPU visit.  Usage: tools/gpu_round.sh TAG [what...]
#   what = tests | tf | engines | bench | ref | ncu | configs | v2 | sanitize | phase        (default: tests bench)
TAG=${1:-r2x}; shift
WHAT=${@:-tests bench}
export PYTHONDONTWRITEBYTECODE=1
mkdir -p gpurun_out
for w in $WHAT; do
  case $w in
    tf) timeout 1500 python -m pytest tests/test_gpu_teacher_forced.py -q > gpurun_out/pytest_tf_$TAG.log 2>&1; echo "tf rc=$?"; tail -15 gpurun_out/pytest_tf_$TAG.log;;
    engines) timeout 1500 python -m pytest tests/test_gpu_engines.py tests/test_gpu_sampler.py -x -q > gpurun_out/pytest_eng_$TAG.log 2>&1; echo "engines rc=$?"; tail -8 gpurun_out/pytest_eng_$TAG.log;;
    tests) timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_$TAG.log;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err;;
    ref) timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json;;
    ncu)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_$TAG python tools/prof_run.py > gpurun_out/prof_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/prof_$TAG.log;;
    plane) timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_leapfrog --launch-skip 3 -c 3 --csv --log-file gpurun_out/plane_leapfrog_$TAG.csv python tools/plane_leapfrog_run.py > gpurun_out/plane_leapfrog_$TAG.log 2>&1; echo "plane rc=$?"; tail -2 gpurun_out/plane_leapfrog_$TAG.log;;
    configs) for c in c2 c3 c4 c5; do timeout 300 python tools/quick.py $c 2>&1 | tail -1; done | tee gpurun_out/configs_$TAG.log;;
    v2) for e in 64,16,107; do NUTS_B200_ENGINE=$e timeout 300 python tools/quick.py c2 2>&1 | tail -1; done | tee gpurun_out/v2_$TAG.log;;
    phase) for c in c2 c5; do NUTS_B200_LIB=$PWD/nuts_rs_b200/libnuts_b200_phase.so timeout 300 python tools/phase_timing.py $c; done 2>&1 | tee gpurun_out/phase_$TAG.log;;
    ncutune) PROF_N=8192 PROF_TUNE_DRAWS=60 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_tune_c5_$TAG python tools/prof_run.py c5 tune > gpurun_out/prof_tune_c5_$TAG.log 2>&1; echo "ncu tune c5 rc=$?"; tail -2 gpurun_out/prof_tune_c5_$TAG.log
             PROF_TUNE_DRAWS=60 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_tune_c2_$TAG python tools/prof_run.py c2 tune > gpurun_out/prof_tune_c2_$TAG.log 2>&1; echo "ncu tune c2 rc=$?"; tail -2 gpurun_out/prof_tune_c2_$TAG.log;;
    ncuc5) timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_c5_$TAG python tools/prof_run.py c5 > gpurun_out/prof_c5_$TAG.log 2>&1; echo "ncu c5 rc=$?"; tail -2 gpurun_out/prof_c5_$TAG.log;;
    user) timeout 900 python -m pytest tests/test_gpu_user_logp.py -q > gpurun_out/pytest_user_$TAG.log 2>&1; echo "user rc=$?"; tail -15 gpurun_out/pytest_user_$TAG.log;;
    ncusample) timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_c2_$TAG python tools/prof_run.py c2 > gpurun_out/prof_c2_$TAG.log 2>&1; echo "ncu c2 rc=$?"; tail -2 gpurun_out/prof_c2_$TAG.log
               timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_chain_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_c4_$TAG python tools/prof_run.py c4 > gpurun_out/prof_c4_$TAG.log 2>&1; echo "ncu c4 rc=$?"; tail -2 gpurun_out/prof_c4_$TAG.log;;
    bsweep) for c in c2 c5 c3; do for b in 1 2 4 8; do NUTS_B200_DRAWS_PER_UNIT=$b timeout 300 python tools/quick.py $c 400 40 3 2>&1 | tail -1 | sed "s/^/B=$b /"; done; done | tee gpurun_out/bsweep_$TAG.log;;
    wide) for e in 64,16,54 128,8,53 128,8,52 256,4,62 256,4,61; do NUTS_B200_ENGINE=$e timeout 300 python tools/quick.py c2 2>&1 | tail -1; done | tee gpurun_out/wide_$TAG.log
          for c in c2 c5; do NUTS_B200_LIB=$PWD/nuts_rs_b200/libnuts_b200_phase.so timeout 300 python tools/quick.py $c 2>&1 | tail -1 | sed "s/^/chunk2 /"; done | tee -a gpurun_out/wide_$TAG.log;;
    variants) for e in 64,16,54 64,16,55 64,16,57 64,16,56 64,16,58 64,16,107; do NUTS_B200_ENGINE=$e timeout 300 python tools/quick.py c2 2>&1 | tail -1; done | tee gpurun_out/variants_$TAG.log;;
    stage) for e in 64,16,54 64,16,58; do NUTS_B200_ENGINE=$e timeout 300 python tools/quick.py c2 2>&1 | tail -1; done | tee gpurun_out/stage_$TAG.log
           NUTS_B200_ENGINE=64,16,58 NUTS_B200_LIB=$PWD/nuts_rs_b200/libnuts_b200_phase.so timeout 300 python tools/phase_timing.py c2 2>&1 | tee -a gpurun_out/stage_$TAG.log
           NUTS_B200_ENGINE=64,16,58 timeout 600 python -m pytest tests/test_gpu_teacher_forced.py -q -k "config2 or checkpoint" 2>&1 | tail -3 | tee -a gpurun_out/stage_$TAG.log;;
    sanitize)
      for c in ${SANITIZE_CASES:-c1 migrate large funnel rank1 cluster}; do
        for tool in racecheck memcheck; do
          NUTS_B200_GRID=3 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $c > gpurun_out/sanitize_${tool}_${c}_$TAG.log 2>&1
          echo "sanitize $tool $c rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok:' gpurun_out/sanitize_${tool}_${c}_$TAG.log | tr '\n' ' ')"
        done
      done;;
  esac
done
