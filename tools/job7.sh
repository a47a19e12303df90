#!/bin/bash
for u in 2 4 8 16; do echo "units target $u"; SSCG_W7_UNITS=$u timeout 120 python tools/w7_time.py 2>&1 | grep "head wgrad"; done
