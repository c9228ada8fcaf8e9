# Bench lines re-taken after profiles/traffic.json was rebuilt from the r02 captures (their roofline
# fields then come from the captures kept beside them), and the coverage workloads with the final code.
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err || echo "bench $tag failed: $(tail -1 gpurun_out/r02_bench_$tag.err)"; }
run rv32i --steps 20
run cov_campaign --workload cov_campaign --steps 5
run cov_campaign_65536 --workload cov_campaign --instances 65536 --steps 5 --no-cpu-baseline
run cov_campaign_veer --workload cov_campaign_veer --steps 5
run veer_class --workload veer_class --steps 10
run veer_class_64k --workload veer_class_64k --steps 10 --no-cpu-baseline
run make_tracing_c --workload make_tracing_c --steps 10 --no-cpu-baseline
run wide_bus --workload wide_bus --steps 10 --no-cpu-baseline
run x_initial --workload x_initial --steps 10 --no-cpu-baseline
run x_initial_il --workload x_initial_il --steps 10 --no-cpu-baseline
bash tools/ncu_one.sh r02_ncu_full_cov_campaign_veer_vlcu_k_run cov_campaign_veer
