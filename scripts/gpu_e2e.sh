#!/bin/bash
# host-pointer (e2e) verify: tests of the pipelined path, then timing against the old single-stage path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or full_size or concurrent" 2>&1 | tail -3
python scripts/e2e_parts.py 2>&1 | tee gpurun_out/e2e_parts.log
