timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/dev_time_net.py 64 fp16 2>&1 | grep "forward\|tailsum\|sum of"
PSSR_V3_NO_PDL=1 timeout 300 python scripts/dev_time_net.py 64 fp16 2>&1 | grep "forward\|tailsum\|sum of"
