#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
python tools/row_bench.py 2>&1 | tail -7
