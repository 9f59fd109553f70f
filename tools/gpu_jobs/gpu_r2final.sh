#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
tail -5 gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
