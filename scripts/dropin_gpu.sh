#!/bin/bash
# Drop-in proof: the byte-identical reference scripts, staged by
# scripts/stage_reference.py into baseline/_ref/, run on the GPU through pointnetgpd_b200.launcher
# against a synthetic $PointNetGPD_FOLDER tree.  Run under gpurun from the repo root; logs -> gpurun_out/.
set -u
ROOTDIR=$(pwd)
REF=$ROOTDIR/baseline/_ref
OUT=$ROOTDIR/gpurun_out
mkdir -p "$OUT"
python scripts/stage_reference.py --verify > "$OUT/r2_dropin_sha256.log" 2>&1 || { echo "staged reference files missing or modified"; cat "$OUT/r2_dropin_sha256.log"; exit 1; }
WORK=$(mktemp -d /tmp/pgpd_dropin.XXXXXX)
mkdir -p "$WORK/PointNetGPD" "$WORK/data"
cp "$REF"/PointNetGPD/main_*.py "$WORK/PointNetGPD/"
cp "$REF"/data/pointnetgpd_3class.model "$WORK/data/"
cd "$WORK/PointNetGPD"
export PYTHONPATH=$ROOTDIR
TREE=$WORK/tree
run() {   # name, timeout, args...
    local name=$1 to=$2; shift 2
    echo "== $name: python -m pointnetgpd_b200.launcher $*" > "$OUT/r2_dropin_$name.log"
    ( time timeout "$to" python -m pointnetgpd_b200.launcher "$@" ) >> "$OUT/r2_dropin_$name.log" 2>&1
    echo "exit code $?" >> "$OUT/r2_dropin_$name.log"
    tail -4 "$OUT/r2_dropin_$name.log" | cut -c1-200
}
run main_1v        600 --synthetic-data "$TREE" main_1v.py --mode train --epoch 1 --batch-size 64 --cuda --gpu 0 --tag dropin1v
ls -la assets/learned_models >> "$OUT/r2_dropin_main_1v.log" 2>&1
# re-load the saved whole-module pickle through the unchanged script's --mode test path (main_1v.py:152-155,184-186)
run main_1v_test   600 --synthetic-data "$TREE" main_1v.py --mode test --batch-size 64 --cuda --gpu 0 --load-model assets/learned_models/dropin1v_0.model
run main_1v_mc     600 --synthetic-data "$TREE" main_1v_mc.py --mode train --epoch 1 --batch-size 64 --cuda --gpu 0 --tag dropinmc
run main_fullv     900 --synthetic-data "$TREE" main_fullv.py --mode train --epoch 1 --batch-size 64 --cuda --gpu 0 --tag dropinfv
run main_test      300 main_test.py --cuda --gpu 0 --load-model ../data/pointnetgpd_3class.model
cd "$ROOTDIR"
rm -rf "$WORK"
