#!/bin/bash
# One gpurun call: quick correctness of the new layout-move paths, then the A/B timings.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== pytest (new paths)"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decoders.py -m gpu -q --timeout=300 \
    -k "gated or layout or cuda_graph or smoke" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"
tail -n 15 gpurun_out/pytest_new.log
echo "== A/B"
timeout 900 python scripts/ab_layout.py 10 > gpurun_out/ab_layout.jsonl 2> gpurun_out/ab_layout.err; echo "rc=$?"
tail -n 5 gpurun_out/ab_layout.err; cat gpurun_out/ab_layout.jsonl
