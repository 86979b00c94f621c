python tools/select_trace.py resnet50 2>&1 | tail -3
python tools/select_trace.py vit_l_16 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/r2_ab.sh tools/probes/libb200prune_old.so 2>&1 | head -4
