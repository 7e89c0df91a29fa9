set -x
( timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 ) 2>&1 | tail -7
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
