#!/bin/bash
for v in "" "PXB_TAYLOR_NBUF=1" "PXB_TAYLOR_GROUPS=2" "PXB_TAYLOR_NBUF=1 PXB_TAYLOR_GROUPS=2"; do
echo "== [$v]"; env $v timeout 300 python tools/profile_stages.py c4 8192 3 2>&1 | grep "propagate"
done
