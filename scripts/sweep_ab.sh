#!/bin/bash
# A/B of library variants on the class sweep (store mode): bash scripts/sweep_ab.sh "" _variant ...
for v in "$@"; do
LB200_LIB_SUFFIX=$v python bench.py --no-fock --no-df3c --no-cpu-baseline --steps 5 --e2e-quartets 65536 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        b=json.loads(ln); p=b['per_class']
        print('variant [$v]', round(b['ms_per_step'],2), {k:round(p[k]['ms'],2) for k in ('2222','2122','2121','2022','2021','2020','1122','1121','1111')})
"
done
