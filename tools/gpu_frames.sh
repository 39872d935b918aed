#!/bin/bash
timeout 600 python -m pytest tests/test_frames.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/frames_bench.py 32
