mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/san_memcheck.log 2>&1; grep -E "ERROR SUMMARY|sanitize run ok" gpurun_out/san_memcheck.log | tail -2
