#!/bin/bash
# Runs tools/run_resident.py with the shipped library and with each kernel variant library found in
# csrc/build/variants (make variant NAME=.. VEXTRA=..).  ADSB_B200_LIB selects the library.
for f in adsbdec_b200/libadsbdec_b200.so adsbdec_b200/csrc/build/variants/lib_*.so; do
  echo "== $f"
  ADSB_B200_LIB=$f timeout 120 python tools/run_resident.py ${1:-60} ${2:-5} 2>&1 | grep -E "wall|^step" | tail -2 | cut -c1-150
done
