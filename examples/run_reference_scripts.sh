#!/bin/bash
# Runs the reference's own example scripts UNCHANGED on the B200 engine (needs a checkout of lcrekko/lq_mpc and a GPU).
# Usage: examples/run_reference_scripts.sh /path/to/lq_mpc
set -e
REF=${1:?path to a checkout of lcrekko/lq_mpc}
HERE=$(cd "$(dirname "$0")/.." && pwd)
python -c 'import sys; sys.path.insert(0, "'$HERE'"); import __graft_entry__ as g; g.build()'
cd "$REF"                                  # error_*.npy / data_lq_mpc_multipleSys.npz are cwd-relative upstream
PYTHONPATH=$HERE/lq_mpc_b200/dropin python working_example_single.py
PYTHONPATH=$HERE/lq_mpc_b200/dropin MPLBACKEND=Agg python working_example_multiple.py
