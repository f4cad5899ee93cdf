#!/bin/bash
DK_PDL=0 timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -6 | cut -c1-200
timeout 300 python tools/kernel_timeline.py --batch 64 2>&1 | tail -6 | cut -c1-200
