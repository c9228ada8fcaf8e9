# Round-end measurements on one B200 (run through gpurun): every bench workload, the reference arm,
# the ncu launch list of the default bench and one `--set full` capture per dominant kernel.
# Results land in gpurun_out/r02_*; copy the ones to keep into profiles/.
set -u
mkdir -p gpurun_out
run() { # <tag> <bench args...>
  tag=$1; shift
  python bench.py "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err || echo "bench $tag failed: $(tail -1 gpurun_out/r02_bench_$tag.err)"
}
run rv32i --steps 20
run rv32i_reference_arm --impl reference --steps 3 --warmup 1
run x_initial --workload x_initial --steps 10
run x_initial_il --workload x_initial_il --steps 10 --no-cpu-baseline
run make_tracing_c --workload make_tracing_c --steps 10
run wide_bus --workload wide_bus --steps 10
run wide_bus_65536 --workload wide_bus --instances 65536 --steps 10 --no-cpu-baseline
run cov_campaign --workload cov_campaign --steps 5
run cov_campaign_65536 --workload cov_campaign --instances 65536 --steps 5 --no-cpu-baseline
run rv32i_16384 --workload rv32i --instances 16384 --steps 10 --no-cpu-baseline
run cov_campaign_mtc --workload cov_campaign_mtc --steps 5
run cov_campaign_veer --workload cov_campaign_veer --steps 5
run veer_class --workload veer_class --steps 10
run veer_class_64k --workload veer_class_64k --steps 10 --no-cpu-baseline
run ports_stream --workload ports_stream --steps 10
run state_stream --workload state_stream --steps 10
tests/widelanes/widelanes 4096 8192 > gpurun_out/r02_widelanes_4096.json
tests/widelanes/widelanes 65536 2048 > gpurun_out/r02_widelanes_65536.json
# launch list of the default bench (kernel shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_rv32i.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-gate > /dev/null 2>&1
for w in rv32i wide_bus cov_campaign cov_campaign_veer make_tracing_c veer_class veer_class_64k x_initial x_initial_il; do
  bash tools/ncu_one.sh r02_ncu_full_${w}_vlcu_k_run $w
done
ls gpurun_out | grep r02_ | wc -l
