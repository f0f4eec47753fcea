#!/bin/bash
# the -m gpu suite alone, with the parity table (element-wise figures) written out
TAG=${1:-t}; OUT=gpurun_out; mkdir -p $OUT
GFB_PARITY_REPORT=$OUT/parity_report_${TAG}.json timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider -rxXs > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" $OUT/pytest_${TAG}.log | tail -5
grep -E "^(FAILED|ERROR)" $OUT/pytest_${TAG}.log | head -20
