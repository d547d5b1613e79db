mkdir -p gpurun_out
(time timeout 200 python -m pytest tests -m gpu -q -k "hme_fullhex or hme_star or hme_golden") > gpurun_out/r02ab_pytest.log 2>&1; tail -12 gpurun_out/r02ab_pytest.log | cut -c1-400
