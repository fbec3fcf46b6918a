mkdir -p gpurun_out/s3
( GLU_SORT_MATCH_FIRST=1 timeout 300 python -m pytest tests/test_sort_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -k "not beyond_2_30" 2>&1 | tail -5 ) > gpurun_out/s3/pytest_mf1.log
cat gpurun_out/s3/pytest_mf1.log
PYTEST_K="reference_cases or skewed" bash tools/exp_sort.sh s3 \
 "GLU_SORT_MATCH_FIRST=0" \
 "GLU_SORT_MATCH_FIRST=1" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CONFIG=1" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CONFIG=6" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CONFIG=8" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CONFIG=4" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CONFIG=0" \
 "GLU_SORT_MATCH_FIRST=0 GLU_SORT_CONFIG=8" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_OPTIONS=2" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CHAIN_ROWS=4" \
 "GLU_SORT_MATCH_FIRST=1 GLU_SORT_CHAIN_ROWS=2"
