set -x
timeout 400 python -m pytest tests -m gpu -x -q --tb=short -k "depth_metrics or batched_tta" 2>&1 | tail -15
