# A/B of two builds of the library on the same box: $1 = alternative .so
L=pruning_for_vision_representation_b200/libb200prune.so
show() { python - "$1" <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1],'value',round(d['value'],2),'ms',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'contract',round(d['roofline']['contract']['ms_per_step'],4))
print('  ', {k: round(v['us'],1) for k,v in d['magnitude']['legs'].items()})
P
}
for rep in 1 2; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_new.json 2>/dev/null; show gpurun_out/ab_new.json
cp $L /tmp/new.so; cp $1 $L
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ab_old.json 2>/dev/null; show gpurun_out/ab_old.json
cp /tmp/new.so $L
done
