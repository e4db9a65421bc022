#!/bin/bash
# run selected gpu tests ($TESTS = pytest -k expression, $FILES = test files) and keep the assertion messages
mkdir -p gpurun_out
timeout 1500 python -m pytest ${FILES:-tests} -q -m gpu -k "$TESTS" --tb=short 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | cut -c1-260 | head -60 > gpurun_out/tests_k.log
cat gpurun_out/tests_k.log
