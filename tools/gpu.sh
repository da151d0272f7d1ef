#!/bin/bash
# One parameterised GPU-box script (run under gpurun):  tools/gpu.sh TASK[,TASK...] [TAG] [-- extra bench args]
#   tests      full `pytest -m gpu` suite (log kept as gpurun_out/pytest_gpu_TAG.log, -rA)
#   multi      tests/test_gpu_multi.py on every visible GPU count the box has (log kept, -rA)
#   smoke      __graft_entry__.smoke()
#   bench      bench.py at N = number of visible GPUs (torchrun for N > 1)  -> gpurun_out/bench_TAG.json
#   bench1     bench.py on one GPU even when the box has more
#   ref        bench.py --impl reference
#   cfgs       bench.py for c1, c2, c3, c5 (one GPU)
#   launches   ncu launch list of a short bench.py run
#   ncu        ncu --set full of the hot kernels on tools/profile_stages.py (KREGEX, PROF_ARGS override)
#   traffic    DRAM bytes of the exchange / Taylor kernels at full c4 and c5 sizes (ncu, 5 metrics) -> gpurun_out/traffic_TAG.csv
#   sanitize   compute-sanitizer memcheck + racecheck over the c1 smoke (and a 2-GPU comb step if 2 GPUs)
#   ab         A/B bench of an environment variant: AB="VAR=value"
TASKS=${1:-tests,bench}
TAG=${2:-run}
shift; shift
[ "$1" = "--" ] && shift
EXTRA="$@"
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | head -3
echo "gpus=$NG cores=$(nproc)"
summ() { python tools/bench_summary.py "$1"; }
run_bench() {  # N out extra...
  local n=$1 out=$2; shift; shift
  if [ "$n" = "1" ]; then
    timeout 1500 python bench.py --gpus 1 "$@" > $out.log 2>&1
  else
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus $n "$@" > $out.log 2>&1
  fi
  grep '^{' $out.log | tail -1 > $out.json
  summ $out.json || tail -30 $out.log
}
for T in ${TASKS//,/ }; do
  echo "== $T"
  case $T in
    tests) timeout 2400 python -m pytest tests -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
           echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu_$TAG.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_$TAG.log | head -20 ;;
    multi) timeout 2400 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_multi_n${NG}_$TAG.log 2>&1
           echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/pytest_gpu_multi_n${NG}_$TAG.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_multi_n${NG}_$TAG.log | head -20 ;;
    smoke) timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ;;
    bench) run_bench $NG gpurun_out/bench_${TAG}_n$NG --steps ${STEPS:-10} --warmup 3 $EXTRA ;;
    bench1) run_bench 1 gpurun_out/bench_${TAG}_n1 --steps ${STEPS:-10} --warmup 3 $EXTRA ;;
    ref) timeout 900 python bench.py --impl reference --steps 1 --warmup 0 $EXTRA 2>&1 | tail -1 | tee gpurun_out/bench_reference_$TAG.json | cut -c1-400 ;;
    cfgs) for c in c1 c2 c3 c5; do echo "-- $c"; run_bench 1 gpurun_out/bench_${TAG}_$c --config $c --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-other-configs $EXTRA; done ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-400} --csv --log-file gpurun_out/launches_$TAG.csv \
                python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs $EXTRA > gpurun_out/bench_ncu_$TAG.log 2>&1; wc -l gpurun_out/launches_$TAG.csv ;;
    ncu) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-exx_eri_kernel|taylor3_kernel|taylor2_kernel|gemm_tma_kernel|theta_kernel|cholqr_kernel|qr_kernel}" \
           -c ${NCU_COUNT:-24} -f -o gpurun_out/prof_$TAG python tools/profile_stages.py ${PROF_ARGS:-c4 2368 1} > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
         python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/ncu_full_$TAG.txt 2>/dev/null; wc -l gpurun_out/ncu_full_$TAG.txt
         # gpurun brings back at most 64 MiB: keep the report itself only when asked to (KEEP_REP=1, few kernels)
         [ "${KEEP_REP:-0}" = "1" ] || rm -f gpurun_out/prof_$TAG.ncu-rep ;;
    traffic) for cfg in "c4 8192" "c5 2048"; do set -- $cfg
               timeout 900 ncu --clock-control none -k regex:"${KREGEX:-exx_eri_kernel|taylor3_kernel|taylor2_kernel}" -c 4 --csv \
                 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum \
                 --log-file gpurun_out/traffic_${TAG}_$1.csv python tools/profile_stages.py $1 $2 1 > gpurun_out/traffic_${TAG}_$1.log 2>&1
               tail -5 gpurun_out/traffic_${TAG}_$1.csv | cut -c1-300; done ;;
    sanitize) for tool in memcheck racecheck; do
                timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
                echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case" gpurun_out/sanitizer_${tool}_$TAG.log | tail -4
              done ;;
    ab) for v in "" "$AB" "" "$AB"; do echo "-- [$v]"; env $v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-other-configs $EXTRA 2>&1 | grep '^{' | tail -1 > gpurun_out/ab.json; summ gpurun_out/ab.json; done ;;
    *) echo "unknown task $T" ;;
  esac
done
