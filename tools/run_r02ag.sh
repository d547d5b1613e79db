mkdir -p gpurun_out
(time timeout 120 python -m pytest tests -m gpu -q -k "aq0_ or shallow_la") > gpurun_out/r02ag_pytest.log 2>&1; tail -12 gpurun_out/r02ag_pytest.log | cut -c1-400
