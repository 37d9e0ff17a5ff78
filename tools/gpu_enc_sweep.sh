#!/bin/bash
export PYTHONUNBUFFERED=1
HF_ENC_TIMING=1 python tools/enc_time.py 2>&1 | grep -E "^op|total" | tail -54
