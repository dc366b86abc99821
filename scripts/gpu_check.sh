#!/bin/bash
# Run on the GPU box via gpurun: GPU test suite, smoke, (optional) sanitizer, short bench.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=300 > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"
tail -n 25 gpurun_out/pytest_kernels.log
echo "== pytest decoders"; timeout 1200 python -m pytest tests/test_gpu_decoders.py -m gpu -q --timeout=600 > gpurun_out/pytest_decoders.log 2>&1; echo "rc=$?"
tail -n 25 gpurun_out/pytest_decoders.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/smoke.log
if [ "${SANITIZE:-0}" = "1" ]; then
  echo "== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/memcheck.log
fi
if [ "${BENCH:-1}" = "1" ]; then
  echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
  tail -n 5 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
