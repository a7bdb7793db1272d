set -x
python -m pytest tests -m gpu -x -q -k "tail_loss or full_network" 2>&1 | tail -5
python scripts/train_prof.py amp 12
