mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -q -k "aq4 or aq5 or aq3 or qg8" ) > gpurun_out/r02t_pytest.log 2>&1; tail -40 gpurun_out/r02t_pytest.log | cut -c1-400
