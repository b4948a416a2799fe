timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "grid or knn or radius" 2>&1 | tail -4
timeout 200 python scripts/knn_batch_lab.py 2>&1 | tail -14

