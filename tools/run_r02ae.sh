mkdir -p gpurun_out
(time timeout 200 python -m pytest tests -m gpu -q -k "golden and (aq4_edge or temporal3)") > gpurun_out/r02ae_pytest.log 2>&1; tail -8 gpurun_out/r02ae_pytest.log | cut -c1-300
(time timeout 100 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -4
