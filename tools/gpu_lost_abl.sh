for f in 0 1 2 3; do echo "dbg=$f"; B200P_LOST_DBG=$f timeout 60 python tools/lost_probe.py 256 20 2; done
