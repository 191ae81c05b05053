#!/bin/bash
# Runs ON THE GPU BOX: parity tests, binary-level end to end with host stage clocks.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2g}
( timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
RBG_HOST_STATS=1 timeout 1200 python tools/e2e_binaries.py --config c2 --reads 10000000 --ref-reads 20000 --out $O/${T}_e2e_binaries.json 2>&1 | cut -c1-900 | tail -12
