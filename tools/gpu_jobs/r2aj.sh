# round-2 GPU job aj: the full GPU suite and smoke() on the last commit
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2aj_pytest.log 2>&1; tail -3 gpurun_out/r2aj_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
