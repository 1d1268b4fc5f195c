#!/bin/bash
# development call: the tests named on the command line, full output kept
TAG=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest "$@" -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -60 gpurun_out/${TAG}_pytest.log
