#!/bin/bash
for dbg in 0 1 2; do
echo "== dbg $dbg NG4"; PXB_TAYLOR_DBG=$dbg timeout 120 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -1
done
echo "== dbg 1 NG2"; PXB_TAYLOR_GROUPS=2 PXB_TAYLOR_DBG=1 timeout 120 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -1
