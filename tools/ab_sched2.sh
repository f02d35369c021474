# A/B of further orderings / priorities of bench.py's step, on top of the "gap" arrangement
run() { python bench.py --no-extra --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' ', round(d['value'],2), round(d['ms_per_step'],4))"; }
for rep in 1 2; do
echo "base";                     run
echo "K1b(cam1) after fwd2";     WSEG_BENCH_S4_START=fwd2 run
echo "fix-up priority -1";       WSEG_BENCH_P5=-1 run
echo "view-2 K1b/bwd priority -1"; WSEG_BENCH_P2=-1 run
done
