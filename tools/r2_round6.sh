python tools/miss_probe.py 2>&1 | tail -24
