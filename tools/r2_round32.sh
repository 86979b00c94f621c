python -m pytest tests/test_gpu_lost.py -m gpu -x -q 2>&1 | tail -2
for cfg in "2 0 128" "1 40 128" "1 60 128" "1 75 128" "1 60 256" "1 50 192" "1 100 128"; do
  set -- $cfg
  echo "== FINISH=$1 SPLIT=$2 THREADS=$3"
  B200P_LOST_FINISH=$1 B200P_LOST_SPLIT_PCT=$2 B200P_LOST_SMALL_THREADS=$3 python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-260
done
