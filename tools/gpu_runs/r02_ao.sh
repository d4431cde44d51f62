#!/bin/bash
# all GPU tests + smoke on the tree with the native trajectory formatter and the device-function unit tests
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -4
