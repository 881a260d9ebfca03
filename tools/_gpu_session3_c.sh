mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest_full.log
tail -4 gpurun_out/s3_pytest_full.log
timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/s3_bench_image_1080p.json 2> gpurun_out/s3_bench_image.err
timeout 120 python tools/bench_image.py 4032 3024 16 > gpurun_out/s3_bench_image_12mp.json 2>> gpurun_out/s3_bench_image.err
cat gpurun_out/s3_bench_image_1080p.json gpurun_out/s3_bench_image_12mp.json; tail -3 gpurun_out/s3_bench_image.err
timeout 300 python bench.py > gpurun_out/s3_bench_default.json 2> gpurun_out/s3_bench_default.err
cut -c1-300 gpurun_out/s3_bench_default.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/s3_launches.csv python bench.py --profile-mode > gpurun_out/s3_ncu.log 2>&1
tail -2 gpurun_out/s3_ncu.log; wc -l gpurun_out/s3_launches.csv
