#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 tools/micro/wbw2 5 800 > $O/wbw2.txt 2>&1; echo "wbw2 rc=$?"; cat $O/wbw2.txt
timeout 900 python -m pytest tests -m gpu -q -k "nonuniform or export or ebal or to_xr or sensitivity" > $O/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_sel.log
