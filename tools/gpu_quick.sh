#!/bin/bash
# quick GPU check: arguments are passed to pytest
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s "$@" > gpurun_out/pytest_quick.log 2>&1
tail -40 gpurun_out/pytest_quick.log
