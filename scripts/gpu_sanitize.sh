#!/bin/bash
# compute-sanitizer over every entry point (memcheck with the big pipelined verify; racecheck and synccheck at small sizes)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_all.py > gpurun_out/r2b_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2b_sanitizer_memcheck.log
SANITIZE_BIG=0 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_all.py > gpurun_out/r2b_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2b_sanitizer_racecheck.log
SANITIZE_BIG=0 timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_all.py > gpurun_out/r2b_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r2b_sanitizer_synccheck.log
