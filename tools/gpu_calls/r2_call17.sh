#!/bin/bash
set -x
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|warn\|return float\|^$\|Docs" | tail -4
bash tools/collect_profiles.sh
