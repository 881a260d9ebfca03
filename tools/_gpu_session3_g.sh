mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py -q > gpurun_out/s3_pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_g.log
tail -5 gpurun_out/s3_pytest_g.log
timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p_dp4a.json 2> gpurun_out/s3_bench_image_g.err
MA_RESAMPLE_DP4A=0 timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p_vec28.json 2>> gpurun_out/s3_bench_image_g.err
timeout 120 python tools/bench_image.py 4032 3024 16 > gpurun_out/s3_bench_image_12mp_dp4a.json 2>> gpurun_out/s3_bench_image_g.err
cat gpurun_out/s3_bench_image_1080p_dp4a.json gpurun_out/s3_bench_image_1080p_vec28.json gpurun_out/s3_bench_image_12mp_dp4a.json
