timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -c 1800 gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
