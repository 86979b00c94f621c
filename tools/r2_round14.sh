python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-300
B200P_LOST_CONV_WARPS=4 python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-300
B200P_LOST_FINISH=0 python tools/lost_probe2.py 256 30 2>&1 | tail -2 | cut -c1-300
python tools/lost_probe2.py 1024 10 2>&1 | tail -1 | cut -c1-300
