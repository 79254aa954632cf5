export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gputest.log 2>&1; echo "gputest rc=$?"
tail -3 gpurun_out/r02g_gputest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02g_smoke.log
timeout 200 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02g_bench.json
