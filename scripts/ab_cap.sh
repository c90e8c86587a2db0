for cfg in "512 0" "1024 0" "2048 0" "512 1000" "1024 1000" "1024 4000"; do
  set -- $cfg
  X="-DRS_DB_QCHUNK=$1"; if [ "$2" != "0" ]; then X="$X -DRS_DB_WAIT_HINT=$2"; fi
  touch rescan_b200/csrc/score.cu
  make -s -j8 -f rescan_b200/csrc/Makefile EXTRA="$X" > /dev/null 2>&1
  echo "QCHUNK $1 HINT $2"
  python scripts/dense_one.py C2 all 2 2>&1 | tail -1
  python scripts/dense_one.py C3 3 2 2>&1 | tail -1
done
