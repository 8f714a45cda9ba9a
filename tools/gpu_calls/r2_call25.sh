#!/bin/bash
set -x
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs\|return float" | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/collect_profiles.sh
