cd "${GRAFT_REPO_ROOT:-/root/repo}"
for T in 32 48; do
timeout 600 python bench.py --no-comparators --no-cpu-baseline --no-extras --triples $T 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('T=$T value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'clk',d['clocks'])"
done
nvidia-smi --query-gpu=memory.used --format=csv
