#!/bin/bash
# compute-sanitizer pass (run under gpurun).  memcheck over the kernel parity tests that exercise the shared-memory
# privatised counters, the TMA / mbarrier pipelines and the warp-private queues; memcheck, racecheck and synccheck
# over tools/sanitize_small.py (one small launch of every kernel family, checked against the oracle -- racecheck is
# ~100x slower than a plain run).  Logs land in gpurun_out/sanitizer_*_<tag>.log; summaries are copied to profiles/.
set -u
TAG=${1:-r2}
SEL='mask_gather or class_encode or resample_confusion_tma or tile_hist or profile_tiles or fit_resize or tap_split'
log=gpurun_out/sanitizer_memcheck_pytest_${TAG}.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "$SEL" -p no:cacheprovider > $log 2>&1
echo "exit code: $?" >> $log
echo "== memcheck over pytest -k '$SEL'"; grep -E "ERROR SUMMARY|passed|failed|exit code" $log | tail -4
for tool in memcheck racecheck synccheck; do
    log=gpurun_out/sanitizer_${tool}_small_${TAG}.log
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_small.py > $log 2>&1
    echo "exit code: $?" >> $log
    echo "== $tool over tools/sanitize_small.py"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small|exit code|Error|hazard" $log | head -8
done
