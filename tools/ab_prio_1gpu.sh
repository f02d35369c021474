# A/B at 1 GPU of the stream priorities of bench.py HotPath.step (P1 view-1 chain, P3 loss chain, P5 sign fix-up, P6 view-2 forward)
for cfg in "-1 -1 0" "-1 -1 -1" "-1 -1 -2" "-2 -1 -1" "-2 -1 -2" "-2 -2 -2"; do set -- $cfg; echo "P1 $1 P3 $2 P5 $3"; WSEG_BENCH_P1=$1 WSEG_BENCH_P3=$2 WSEG_BENCH_P5=$3 python bench.py --no-extra --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' ', round(d['value'],2), round(d['ms_per_step'],4))"; done
