set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/s3_gpu.txt
timeout 600 python -m pytest tests/test_pdl_gpu.py tests/test_image_gpu.py -x -q > gpurun_out/s3_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_new.log
for i in 1 2; do
MA_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench_pdl0_$i.json 2> gpurun_out/s3_bench_pdl0_$i.err
MA_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench_pdl1_$i.json 2> gpurun_out/s3_bench_pdl1_$i.err
done
MA_PDL=1 MA_ENCODER_STREAMS=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s3_bench_pdl1_s1.json 2> gpurun_out/s3_bench_pdl1_s1.err
MA_PDL=1 timeout 300 python bench.py --steps 5 --warmup 3 --views 100 > gpurun_out/s3_bench_v100_pdl1.json 2> gpurun_out/s3_bench_v100_pdl1.err
MA_PDL=0 timeout 300 python bench.py --steps 5 --warmup 3 --views 100 > gpurun_out/s3_bench_v100_pdl0.json 2> gpurun_out/s3_bench_v100_pdl0.err
timeout 400 python tools/scale_check.py 1000 > gpurun_out/s3_scale1000.log 2>&1
tail -3 gpurun_out/s3_pytest_new.log; cat gpurun_out/s3_bench_pdl*.json | cut -c1-200; cat gpurun_out/s3_scale1000.log | tail -3
