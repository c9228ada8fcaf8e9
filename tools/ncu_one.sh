# usage: tools/ncu_one.sh <tag> <workload> [VLCU_LIB]   -> gpurun_out/<tag>.csv
# one launch of the workload's vlcu_k_run under `ncu --set full --clock-control none`, as the compact
# metric,unit,value table kept under profiles/ (tools/ncu_compact.py)
tag=$1; wl=$2; [ -n "$3" ] && export VLCU_LIB=$3
ncu --set full --clock-control none -k regex:vlcu_k_ -c 1 -s 4 -o /tmp/$tag python bench.py --workload $wl --steps 2 --no-cpu-baseline --no-parity-gate $BENCH_ARGS > /dev/null 2>/tmp/$tag.err
tail -1 /tmp/$tag.err
ncu -i /tmp/$tag.ncu-rep --page raw --csv > /tmp/$tag.raw.csv; python3 tools/ncu_compact.py /tmp/$tag.raw.csv gpurun_out/$tag.csv
