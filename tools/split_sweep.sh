run() { timeout 120 python tools/run_resident.py 60 8 2>&1 | grep -E "wall" | tail -5 | awk '{print $2}' | tr '\n' ' '; echo; }
for split in none 30,30,30,10 32,32,32,4 45,45,10 25,25,25,25 34,33,23,10; do
  echo -n "== JOB_SPLIT=$split: "
  if [ "$split" = "none" ]; then run; else ADSB_B200_JOB_SPLIT=$split run; fi
done
