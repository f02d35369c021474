# A/B at 2 GPUs of the loss chain's stream priority (bench.py HotPath.step): the loss chain ends the step once prototypes travel
for cfg in "-1 -1" "-1 -2" "0 -2"; do set -- $cfg; echo "P1 $1 P3 $2"; WSEG_BENCH_P1=$1 WSEG_BENCH_P3=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' ', round(d['value'],2), round(d['ms_per_step'],4))"; done
