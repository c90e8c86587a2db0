for cfg in "512 4 8" "384 5 8" "512 6 4"; do
  set -- $cfg
  make -s -j8 -f rescan_b200/csrc/Makefile EXTRA="-DRS_DB_CAP=$1 -DRS_DB_BPS=$2 -DRS_DB_WARPS=$3" > /dev/null 2>&1
  touch rescan_b200/csrc/score.cu
  echo "CAP $1 BPS $2 WARPS $3"
  python scripts/dense_one.py C2 all 2 2>&1 | tail -1
  python scripts/dense_one.py C3 3 2 2>&1 | tail -1
done
