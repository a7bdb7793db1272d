set -x
python -m pytest tests -m gpu -x -q -k "graphed_predictor" 2>&1 | tail -40
