#!/bin/bash
tag=${1:-s12}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "sigma" > gpurun_out/${tag}_tests_sigma.log 2>&1; tail -8 gpurun_out/${tag}_tests_sigma.log
