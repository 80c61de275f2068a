#!/bin/bash
# last GPU call of round 2: device assembly tests, the GPU suite (minus its one long test), the realization loop
mkdir -p gpurun_out
( timeout 90 python -m pytest tests/test_device_assembly.py -m gpu -x -q --durations=3; echo "asm tests rc=$?" ) > gpurun_out/r2g_asm_tests.log 2>&1
tail -3 gpurun_out/r2g_asm_tests.log
( timeout 140 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity_large.py::test_recycling_solvers_1001_128 --durations=5; echo "suite rc=$?" ) > gpurun_out/r2g_gputests.log 2>&1
tail -3 gpurun_out/r2g_gputests.log
( timeout 80 python scripts/bench_realizations.py 501 64; echo "realizations rc=$?" ) > gpurun_out/r2g_bench_realizations_501_64.log 2>&1
tail -6 gpurun_out/r2g_bench_realizations_501_64.log | cut -c1-400
