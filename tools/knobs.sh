#!/bin/bash
# Step time of the resident bench under the development knobs of api.cu.
for ts in 0 1; do for sms in 8 16; do
  echo "== ADSB_B200_TAIL_SPLIT=$ts ADSB_B200_RESOLVE_SMS=$sms"
  ADSB_B200_TAIL_SPLIT=$ts ADSB_B200_RESOLVE_SMS=$sms timeout 120 python tools/run_resident.py 60 6 2>&1 | tail -4 | grep -E "wall|step" | tail -2 | cut -c1-120
done; done
