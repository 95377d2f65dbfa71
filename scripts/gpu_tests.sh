#!/bin/bash
OUT=gpurun_out/${1:-tests}; mkdir -p $OUT
( time timeout 2400 python -m pytest tests -m gpu -q ${PYTEST_ARGS} ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
