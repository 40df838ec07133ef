#!/bin/bash
# Round 2, the GPU suite on the round's last source.
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $out/r02y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/r02y_pytest.log
