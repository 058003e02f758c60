#!/bin/bash
# A/B of layer-3 kernel variants (train step), then the full GPU test suite on the candidate build $CAND
mkdir -p gpurun_out
PAT="k_l3_fwd|eager fwd" bash scripts/gpu_ab.sh $VARIANTS
if [ -n "$CAND" ]; then
  cp pointnetgpd_b200/libpgpd.so /tmp/libpgpd_saved.so
  cp build/variants/libpgpd_$CAND.so pointnetgpd_b200/libpgpd.so
  timeout 600 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -5
  cp /tmp/libpgpd_saved.so pointnetgpd_b200/libpgpd.so
fi
