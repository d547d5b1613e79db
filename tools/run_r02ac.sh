mkdir -p gpurun_out
(time timeout 200 python -m pytest tests -m gpu -q -k "aq4_tl3") > gpurun_out/r02ac_pytest.log 2>&1; tail -12 gpurun_out/r02ac_pytest.log | cut -c1-400
