mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/scale_check.py 1000 > gpurun_out/s3_scale1000_n8.log 2> gpurun_out/s3_scale1000_n8.err
echo "rc=$?"; cat gpurun_out/s3_scale1000_n8.log; tail -5 gpurun_out/s3_scale1000_n8.err
