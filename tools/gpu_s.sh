#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -4 gpurun_out/s_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
