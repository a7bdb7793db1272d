python -m pytest tests -m gpu -x -q -k "graphed_predictor or optimize_for_inference or merge_bias" 2>&1 | grep -v "^$" | tail -45
