#!/bin/bash
# build + CPU-side checks to run before every gpurun call
set -e
cd "$(dirname "$0")/.."
python noise-robust-vit_b200/build.py
python -m pytest tests/test_abi.py -q -x 2>&1 | tail -2
