for B in 256; do
for F in 1 2; do
echo "FINISH=$F"; B200P_LOST_FINISH=$F timeout 120 python tools/lost_probe2.py $B 30 2>&1 | tail -2 | cut -c1-150
done; done
