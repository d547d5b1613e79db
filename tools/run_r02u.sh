mkdir -p gpurun_out
(time timeout 700 python -m pytest tests -m gpu -q -k "temporal or lookahead_threads" ) > gpurun_out/r02u_pytest.log 2>&1; tail -40 gpurun_out/r02u_pytest.log | cut -c1-400
