mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -q -k "12 and not 1280 and not cfg") > gpurun_out/r02ad_pytest.log 2>&1; tail -25 gpurun_out/r02ad_pytest.log | cut -c1-400
