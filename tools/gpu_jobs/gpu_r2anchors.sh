#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_control_sequences.py -m gpu -q -k "anchors" > gpurun_out/r2anchors.log 2>&1; tail -4 gpurun_out/r2anchors.log; grep -n "^E " gpurun_out/r2anchors.log | head -10
