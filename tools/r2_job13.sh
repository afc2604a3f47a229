#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_parity.json
( timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) > gpurun_out/j13_pytest.log
cat gpurun_out/j13_pytest.log
