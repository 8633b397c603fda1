#!/bin/bash
# Step time of the resident bench under the development knobs of api.cu.
for sms in ${SMS_LIST:-2 4 8 12 16}; do
  echo "== ADSB_B200_RESOLVE_SMS=$sms"
  ADSB_B200_RESOLVE_SMS=$sms timeout 120 python tools/run_resident.py 60 6 2>&1 | tail -4 | grep -E "wall|step" | tail -2 | cut -c1-150
done
