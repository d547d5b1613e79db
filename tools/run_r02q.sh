mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -q -k "hist or temporal or fadein or fades_10fps") > gpurun_out/r02q_pytest.log 2>&1; tail -30 gpurun_out/r02q_pytest.log | cut -c1-250
