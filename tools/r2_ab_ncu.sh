L=pruning_for_vision_representation_b200/libb200prune.so
M=gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,sm__cycles_elapsed.max,smsp__inst_executed.sum
python tools/select_probe.py resnet50 2 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k "regex:k_select_bracket|k_select_sample|k_select_finish" -s 6 -c 6 --csv --log-file gpurun_out/ab_ncu_new.csv python tools/select_probe.py resnet50 2 > /dev/null 2>&1
cp $L /tmp/new.so; cp $1 $L
timeout 300 ncu --metrics $M --clock-control none -k "regex:k_select_bracket|k_select_sample|k_select_finish" -s 6 -c 6 --csv --log-file gpurun_out/ab_ncu_old.csv python tools/select_probe.py resnet50 2 > /dev/null 2>&1
cp /tmp/new.so $L
python - <<'P'
import csv
for tag in ('old','new'):
    rows=[r for r in csv.reader(open(f'gpurun_out/ab_ncu_{tag}.csv')) if len(r)>10]
    h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
    d={}
    for r in rows[1:]:
        d.setdefault((r[ii],r[ki][:28]),{})[r[mi]]=r[vi]
    for k,v in d.items(): print(tag,k,{a.replace('sm__cycles_','').replace('gpu__time_',''):b for a,b in v.items()})
P
