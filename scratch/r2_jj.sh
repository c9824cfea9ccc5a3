#!/bin/bash
# call JJ: full GPU suite after the slab initial conditions / slab driver / session changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/jj_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/jj_pytest.log
