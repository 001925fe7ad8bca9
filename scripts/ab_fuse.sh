python scripts/bench_block.py
for m in 0 1 2 4 6 7; do echo "mask $m: $(SSD3D_FUSE_DWPW=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])')"; done
