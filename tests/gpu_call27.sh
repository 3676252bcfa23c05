#!/bin/bash
mkdir -p gpurun_out
python -c "import torch" > /dev/null 2>&1
which compute-sanitizer
timeout -s KILL 400 compute-sanitizer --tool memcheck --target-processes all python tests/gpu_diag.py enc-micro > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|enc micro" gpurun_out/sanitizer_memcheck.log | head -12
timeout -s KILL 400 compute-sanitizer --tool racecheck --target-processes all python tests/gpu_diag.py enc-micro > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|hazard|enc micro" gpurun_out/sanitizer_racecheck.log | head -12
timeout -s KILL 300 compute-sanitizer --tool synccheck --target-processes all python tests/gpu_diag.py mel > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"; grep -E "ERROR SUMMARY|Barrier error|mel n_mels=128 tone" gpurun_out/sanitizer_synccheck.log | head -8
