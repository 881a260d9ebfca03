mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py tests/test_model_gpu.py::test_forward_tiny_info_sharing_variants tests/test_geometric_gpu.py -q > gpurun_out/s3_pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_e.log
tail -6 gpurun_out/s3_pytest_e.log
timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p_c.json 2> gpurun_out/s3_bench_image_c.err
MA_RESAMPLE_BYTE_LOADS=1 timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p_bytes.json 2>> gpurun_out/s3_bench_image_c.err
timeout 120 python tools/bench_image.py 4032 3024 16 > gpurun_out/s3_bench_image_12mp_c.json 2>> gpurun_out/s3_bench_image_c.err
cat gpurun_out/s3_bench_image_1080p_c.json gpurun_out/s3_bench_image_1080p_bytes.json gpurun_out/s3_bench_image_12mp_c.json
