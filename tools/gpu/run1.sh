mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/a_tests.log
cat gpurun_out/a_tests.log
