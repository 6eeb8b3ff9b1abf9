#!/usr/bin/env bash
# compute-sanitizer over the smoke path (these kernels hand-roll mbarrier rings, release / acquire
# counters in L2 and fp64 atomics).  ONE tool per invocation and per GPU-box call (profiling guide: several tools in one
# call have left GPUs unusable):
#
#     tools/sanitize.sh memcheck  [precision] [small|teacher]     # out-of-bounds / misaligned global + shared accesses
#     tools/sanitize.sh racecheck [precision] [small|teacher]     # shared-memory hazards between barriers
#     tools/sanitize.sh synccheck [precision] [small|teacher]     # invalid barrier / mbarrier usage
#
# The plain run goes first (a program that faults must not run under the tool); the log lands in gpurun_out/ and, when
# it is worth keeping, is copied to profiles/.
set -u
tool="${1:-memcheck}"
prec="${2:-fp16}"
tag="${3:-small}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
log="gpurun_out/sanitizer_${tool}_${prec}_${tag}.txt"
python tools/sanitize_smoke.py "$prec" "$tag" > "gpurun_out/sanitize_plain.log" 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
extra=""
[ "$tool" = "racecheck" ] && extra="--racecheck-report all"
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool "$tool" $extra --print-limit 40 --log-file "$log" \
    python tools/sanitize_smoke.py "$prec" "$tag" > "gpurun_out/sanitize_run.log" 2>&1
rc=$?
tail -3 gpurun_out/sanitize_run.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" "$log" | sort | uniq -c | sort -rn | head -20
exit $rc
