# round 2, call 1: LOST count-only path
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lost.py -m gpu -q -x --timeout 300 ) > gpurun_out/r2_pytest_lost.log 2>&1; echo "pytest lost rc=$?"; tail -15 gpurun_out/r2_pytest_lost.log
python tools/lost_probe2.py 256 20 2>&1 | tail -2
B200P_LOST_CONV_WARPS=4 python tools/lost_probe2.py 256 20 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_lost -c 24 --csv --log-file gpurun_out/r2_lost_launches.csv python tools/lost_probe2.py 256 2 > gpurun_out/r2_ncu_lost.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_lost_launches.csv')) if len(r)>5]
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    print(d['ID'], d['Kernel Name'][:60], d['Metric Name'], d['Metric Value'])
PY
