#!/bin/bash
# mlp_bwd switch experiments (A/B library built with EXTRA=-DBD_BWD_DBG): per-kernel ms at 2^17 rows
export BD_B200_LIB=$PWD/big_dreamer_b200/libbd_b200_dbg.so
for P in 0 1; do for V in 0 1 2 3; do
BD_BWD_PAIR=$P BD_BWD_DBGV=$V timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 5 --warmup 3 --rows 131072 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pair=$P dbg=$V', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
"
done; done
