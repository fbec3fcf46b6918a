mkdir -p gpurun_out/s4
( timeout 300 python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "host or reference_cases" 2>&1 | tail -5 ) > gpurun_out/s4/pytest.log
cat gpurun_out/s4/pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/s4/bench.log
cat gpurun_out/s4/bench.log
