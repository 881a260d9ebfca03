mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_image_gpu.py::test_preprocess_inputs_feeds_infer -q > gpurun_out/s3_pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_f.log
tail -4 gpurun_out/s3_pytest_f.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:resample -c 6 -o gpurun_out/s3_image_kernels python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_ncu_image.log 2>&1
MA_RESAMPLE_BYTE_LOADS=1 timeout 200 ncu --set full --import-source on --clock-control none -k regex:resample_h -c 3 -o gpurun_out/s3_image_kernels_bytes python tools/bench_image.py 1920 1080 64 >> gpurun_out/s3_ncu_image.log 2>&1
tail -3 gpurun_out/s3_ncu_image.log; ls -la gpurun_out/*.ncu-rep | tail -3
