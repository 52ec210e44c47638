# final state of round 2: repeatability of the e2e figures, the whole GPU suite, smoke()
timeout 120 python tools/ml_diag.py > gpurun_out/ml_diag.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
cat gpurun_out/ml_diag.txt; cat gpurun_out/pytest.txt; tail -n 3 gpurun_out/smoke.txt
