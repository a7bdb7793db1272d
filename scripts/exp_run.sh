python -m pytest tests -m gpu -x -q --tb=short 2>&1 | grep -v "^$" | tail -30
