#!/bin/bash
# Step time of the resident bench under the development knobs of api.cu, for the shipped library and the variants.
run() { timeout 120 python tools/run_resident.py 60 8 2>&1 | grep -E "wall|^step" | tail -6 | awk '/wall/{printf "%s ", $2} /^step/{f=$6" "$12} END{print " fused/gate+slice " f}'; }
for lib in adsbdec_b200/libadsbdec_b200.so adsbdec_b200/csrc/build/variants/lib_*.so; do
  [ -f "$lib" ] || continue
  for sms in ${SMS_LIST:-0 4 8}; do
    for split in ${SPLIT_LIST:-none 20,30,30,20}; do
      echo -n "== $(basename $lib) RESOLVE_SMS=$sms JOB_SPLIT=$split: "
      if [ "$split" = "none" ]; then ADSB_B200_LIB=$lib ADSB_B200_RESOLVE_SMS=$sms run; else ADSB_B200_LIB=$lib ADSB_B200_RESOLVE_SMS=$sms ADSB_B200_JOB_SPLIT=$split run; fi
    done
  done
done
