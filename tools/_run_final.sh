# final state of round 2: the whole GPU suite, smoke()
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
cat gpurun_out/pytest.txt; tail -n 3 gpurun_out/smoke.txt
