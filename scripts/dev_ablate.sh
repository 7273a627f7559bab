PSSR_V3_SMALL=1 PSSR_V3_VERBOSE=1 timeout 300 python scripts/dev_time_net.py 64 fp16 2>&1 | grep "16x16\|8x8\|forward"
