# A/B of the step's stream arrangement (bench.py HotPath.step): when view 2's K1b + backward may start, and their priority
for cfg in "proto2 0" "fwd2 0" "fwd2 -1" "fwd2 -2" "proto2 -1"; do set -- $cfg; echo "s2 after $1, priority $2"; WSEG_BENCH_S2_AFTER=$1 WSEG_BENCH_P2=$2 python bench.py --no-extra --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' ', round(d['value'],2), round(d['ms_per_step'],4))"; done
