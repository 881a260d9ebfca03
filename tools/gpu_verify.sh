#!/bin/bash
# Round-end verification on a B200 box (run through gpurun): full GPU test suite, the default bench line (with the CPU and
# GPU-eager baselines), the reference arm, the ncu launch list of one step.  Everything lands in gpurun_out/ with the given tag.
TAG=${1:-verify}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-260 gpurun_out/${TAG}_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-260 gpurun_out/${TAG}_bench_reference.json
python bench.py --profile-mode > gpurun_out/${TAG}_plain_profile.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-mode > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
