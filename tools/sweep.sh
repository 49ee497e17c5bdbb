#!/bin/bash
# BASELINE.json configs[4]: key-size sweep (encrypt throughput + IMAD-roofline fraction, CPU port beside it) and configs[3] per GPU.
# Run on a B200 box: bash tools/sweep.sh > gpurun_out/sweep.jsonl
set -e
cd "$(dirname "$0")/.."
python bench.py --n-bits 1024 --units 262144 --steps 3 --warmup 3 --no-witness
python bench.py --n-bits 2048 --units 65536 --steps 3 --warmup 3 --no-witness
python bench.py --n-bits 3072 --units 32768 --steps 3 --warmup 3 --no-witness
python bench.py --n-bits 4096 --units 16384 --steps 3 --warmup 3 --no-witness
python bench.py --workload witness --n-bits 3072 --units 32768 --steps 2 --warmup 2
python bench.py --workload witness --n-bits 2048 --units 65536 --steps 2 --warmup 2
python bench.py --workload tally --steps 5 --warmup 3
