for cfg in "1 100 512" "2 0 128" "1 50 192" "1 30 256" "1 25 192" "1 35 160" "1 100 512" "1 15 256"; do
  set -- $cfg
  echo "== FINISH=$1 SPLIT=$2 THREADS=$3"
  B200P_LOST_FINISH=$1 B200P_LOST_SPLIT_PCT=$2 B200P_LOST_SMALL_THREADS=$3 python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-130
done
