#!/bin/bash
# compute-sanitizer over the tiny configuration (SURVEY.md §5 "race detection"): memcheck of smoke() (one
# forward + backward of the TINY denoiser: every kernel family incl. the CTA-pair GEMM, attention fwd/bwd, the second-stream weight
# gradients) and racecheck (shared-memory hazards) of the attention / GEMM kernel tests at their smallest shapes.
# usage: tools/sanitize.sh [outdir]      -> <outdir>/sanitize_memcheck.log, sanitize_racecheck.log, sanitize_summary.txt
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
export OF_SANITIZE=1
timeout ${SANITIZER_TIMEOUT:-900} $SAN --tool memcheck --print-limit 20 --error-exitcode 9 \
    python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitize_memcheck.log" 2>&1
echo "memcheck rc $?" >> "$OUT/sanitize_memcheck.log"
timeout ${SANITIZER_TIMEOUT:-900} $SAN --tool racecheck --racecheck-report analysis --print-limit 20 --error-exitcode 9 \
    python -m pytest -x -q tests/test_kernels_gpu.py -k "(test_gemm_conv_forward and 2-200-96-96-3) or (test_gemm_dgrad_mn_major and 2-200-136-96-3) or (test_gemm_wgrad_splitk and 2-200-96-136-3) or (test_attention_forward and 1-128-1-1-64 and 0) or (test_attention_backward and 1-128-1-1-64)" \
    > "$OUT/sanitize_racecheck.log" 2>&1
echo "racecheck rc $?" >> "$OUT/sanitize_racecheck.log"
{
  echo "== memcheck: smoke() on the TINY config"; grep -E "ERROR SUMMARY|memcheck rc|smoke:" "$OUT/sanitize_memcheck.log" | tail -5
  echo "== racecheck: kernel tests (GEMM forward, attention fwd/bwd at B1 L128 H1 D64)"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|racecheck rc|passed|failed" "$OUT/sanitize_racecheck.log" | tail -6
} > "$OUT/sanitize_summary.txt"
cat "$OUT/sanitize_summary.txt"
