#!/bin/bash
# Runs ON THE GPU BOX: what the driver runs at round end -- smoke(), the whole -m gpu suite, the default bench -- on the final sources.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-r2u}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${T}_smoke.log
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee $O/${T}_pytest.log
timeout 1200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -20 $O/${T}_bench.err
python tools/bench_summary.py $O/${T}_bench.json
