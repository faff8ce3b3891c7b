#!/bin/bash
# GPU test run with a per-test time limit and a streamed, verbose log (a hanging kernel must not eat the GPU budget):
#   tools/gpu_tests.sh <log> [pytest args...]
LOG=$1; shift
timeout 1200 python -m pytest "$@" -m gpu -x -v --timeout=150 --timeout-method=thread > "$LOG" 2>&1
echo "pytest rc=$?" >> "$LOG"
