mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest_full2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_full2.log
tail -6 gpurun_out/s3_pytest_full2.log
timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p_b.json 2> gpurun_out/s3_bench_image_b.err
MA_RESAMPLE_ONE_ROW=1 timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p_onerow.json 2>> gpurun_out/s3_bench_image_b.err
timeout 120 python tools/bench_image.py 4032 3024 16 > gpurun_out/s3_bench_image_12mp_b.json 2>> gpurun_out/s3_bench_image_b.err
cat gpurun_out/s3_bench_image_1080p_b.json gpurun_out/s3_bench_image_1080p_onerow.json gpurun_out/s3_bench_image_12mp_b.json
