# usage: tools/ncu_src.sh <tag> <workload> [VLCU_LIB]   -> gpurun_out/<tag>.csv (compact metrics) and
# gpurun_out/<tag>.src.csv.gz (ncu --page source --csv of the same launch: per-SASS-line counts)
tag=$1; wl=$2; [ -n "$3" ] && export VLCU_LIB=$3
ncu --set full --clock-control none --import-source on -k regex:vlcu_k_ -c 1 -s 4 -o /tmp/$tag python bench.py --workload $wl --steps 2 --no-cpu-baseline --no-parity-gate $BENCH_ARGS > /dev/null 2>/tmp/$tag.err
tail -1 /tmp/$tag.err
ncu -i /tmp/$tag.ncu-rep --page raw --csv > /tmp/$tag.raw.csv; python3 tools/ncu_compact.py /tmp/$tag.raw.csv gpurun_out/$tag.csv
ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/$tag.src.csv.gz
ls -la gpurun_out/$tag.*
