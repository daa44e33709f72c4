#!/bin/bash
echo "== impl 2"; S2L_TC_IMPL=2 timeout 300 python tools/tc_experiments.py --child
S2L_TC_IMPL=2 timeout 600 python -m pytest tests -q -m gpu -x -k "tc or cta or vol or plain or precision" 2>&1 | tail -3
