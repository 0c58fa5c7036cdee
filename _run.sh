python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --batch 32768 --steps 2 --warmup 3 --no-cpu --phases 2>&1 | grep -v "^{" | tr ',' '\n' | grep "sort\|total"
python bench.py --batch 32768 --steps 3 --warmup 3 --no-cpu 2>&1 | grep -o '"value": [0-9.]*' | head -1
