set -e
timeout 300 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "staged or simpo_backward or segment" 2>&1 | tail -3
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['value'])"; }
OSPO_HEAD_OVERLAP=2 run 29511 two
run 29512 three
OSPO_HEAD_OVERLAP=2 run 29513 two
run 29514 three
