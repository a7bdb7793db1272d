#!/bin/sh
# Copies the files of the UNMODIFIED reference that the hot path needs from the read-only checkout into
# baseline/_ref/ -- git-ignored (the reference's sources are never committed here) but shipped to the GPU box by gpurun, where
# /root/reference does not exist.  Run by __graft_entry__.build() whenever /root/reference is present.
#   baseline/_ref/code/models/{__init__,model,sadecoder,losses,get_network,get_lossfunc}.py   the reference modules
#   baseline/_ref/code/creator/*.py                                                        optimizer / parameter-group factories
#   baseline/_ref/code/utils/eval_utils.py                                                 inference helpers (TTA, metrics)
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
[ -d "$SRC/code/models" ] || { echo "fetch_ref: $SRC/code/models not found (nothing to do)"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST/code/models" "$DST/code/creator" "$DST/code/utils"
for f in __init__ model sadecoder losses get_network get_lossfunc; do cp "$SRC/code/models/$f.py" "$DST/code/models/"; done
cp "$SRC"/code/creator/*.py "$DST/code/creator/"
cp "$SRC/code/utils/eval_utils.py" "$DST/code/utils/"
cp "$SRC/LICENSE" "$DST/" 2>/dev/null || true
echo "fetch_ref: reference modules copied to $DST"
