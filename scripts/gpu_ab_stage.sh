# A/B at N = 2: backward in two parts (default) against three (dX beside the second all-reduce, OSPO_HEAD_OVERLAP=3)
set -e
timeout 300 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "staged or simpo_backward or segment" 2>&1 | tail -3
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['value'])"; }
run 29511 two
OSPO_HEAD_OVERLAP=3 run 29512 three
run 29513 two
OSPO_HEAD_OVERLAP=3 run 29514 three
