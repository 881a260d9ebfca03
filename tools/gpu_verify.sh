#!/bin/bash
# Round-end verification on a B200 box (run through gpurun): full GPU test suite, the default bench line, the input-side
# micro-benchmark and the ncu launch list of one step.  Everything lands in gpurun_out/ with the given tag.
TAG=${1:-verify}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-260 gpurun_out/${TAG}_bench.json
MA_DPT_CHUNK=8 timeout 300 python bench.py > gpurun_out/${TAG}_bench_dpt8.json 2> gpurun_out/${TAG}_bench_dpt8.err
cut -c1-260 gpurun_out/${TAG}_bench_dpt8.json
timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/${TAG}_image_1080p.json 2> gpurun_out/${TAG}_image.err
MA_RESAMPLE_V_ROLLED=1 timeout 120 python tools/bench_image.py 1920 1080 64 > gpurun_out/${TAG}_image_1080p_vrolled.json 2>> gpurun_out/${TAG}_image.err
timeout 120 python tools/bench_image.py 4032 3024 16 > gpurun_out/${TAG}_image_12mp.json 2>> gpurun_out/${TAG}_image.err
cat gpurun_out/${TAG}_image_1080p.json gpurun_out/${TAG}_image_1080p_vrolled.json gpurun_out/${TAG}_image_12mp.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-mode > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
