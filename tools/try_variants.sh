#!/bin/bash
# Runs tools/run_resident.py with each kernel variant library found in csrc/build/variants.
cp adsbdec_b200/libadsbdec_b200.so /tmp/lib_backup.so
for f in adsbdec_b200/csrc/build/variants/lib_*.so; do
  cp $f adsbdec_b200/libadsbdec_b200.so
  echo "== $f"
  timeout 90 python tools/run_resident.py ${1:-60} 4 2>&1 | grep "^step" | tail -2
done
cp /tmp/lib_backup.so adsbdec_b200/libadsbdec_b200.so
