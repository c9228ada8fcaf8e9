# usage: tools/ab_veer.sh "<variants>" "<workload[:instances]> ..."   (variant "default" = the in-tree library)
python -m pytest tests/test_veer_class_gpu.py -x -q 2>&1 | tail -1
for v in $1; do
  for w in $2; do
    wl=${w%%:*}; n=""; [ "$w" != "$wl" ] && n="--instances ${w#*:}"
    if [ $v = default ]; then L=""; else L="VLCU_LIB=$PWD/verilator_b200/libvlcu_exp_veer_$v.so"; fi
    env $L python bench.py --workload $wl $n --steps 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v','$w', round(d['value']/1e6,1),'M/s', round(d['ms_per_step'],1),'ms')"
  done
done
