# Multi-GPU lines on ONE box (run through `gpurun --gpus 8`): the default workload and the coverage
# campaign under torchrun at N = 8 (and N = 2, 4 when asked: NS="2 4 8"), e2e in seeds mode (the
# default) and, for comparison, with host program images.  Results land in gpurun_out/r02_*_<N>gpu.json.
set -u
mkdir -p gpurun_out
NS=${NS:-8}
port=29500
tr() { # <n> <tag> <bench args...>
  n=$1; tag=$2; shift 2; port=$((port + 1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n "$@" > gpurun_out/r02_bench_${tag}_${n}gpu.json 2> gpurun_out/r02_bench_${tag}_${n}gpu.err \
    || echo "bench $tag x$n failed: $(tail -2 gpurun_out/r02_bench_${tag}_${n}gpu.err)"
}
for n in $NS; do
  tr $n rv32i --steps 10 --no-cpu-baseline
  tr $n cov_campaign --workload cov_campaign --steps 5 --no-cpu-baseline
done
n=$(echo $NS | awk '{print $NF}')
tr $n rv32i_images --steps 10 --no-cpu-baseline --e2e-mode images
tr $n veer_class --workload veer_class --steps 5 --no-cpu-baseline
tr $n rv32i_reference_arm --impl reference --steps 2 --warmup 1
ls gpurun_out | grep gpu.json
