N=${1:-8}
run() {  # run <tag> [env...]
  tag=$1; shift
  echo "=== bench $N GPUs [$tag]"
  env "$@" BENCH_BREAKDOWN=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench$N.$tag.log 2>&1
  rc=$?
  grep -E "Error|error:|Traceback|raise |File \"/root|File \"/tmp|CUDA|NCCL WARN|trap|assert" gpurun_out/bench$N.$tag.log | grep -v "errors.html\|error_file\|NCCL version" | head -24 | cut -c1-260
  grep -E "^\{\"metric|BREAKDOWN" gpurun_out/bench$N.$tag.log | cut -c1-1700
  return $rc
}
run default X=1 || run nvls_off TRLX_B200_NVLS=0 || run overlap_off TRLX_B200_NVLS=0 TRLX_B200_OVERLAP_GRAD=0
