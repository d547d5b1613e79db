mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -q -k "hme or base8 or pool16_10bit" ) > gpurun_out/r02s_pytest.log 2>&1; tail -40 gpurun_out/r02s_pytest.log | cut -c1-300
