#!/bin/bash
for v in 10 11; do for z in 1 2 3; do echo -n "variant $v ntz $z: "; XINV_FUSED3_VARIANT=$v XINV_FUSED3_NTZ=$z timeout 120 python scripts/prof_c3.py 200 | tail -1; done; done
for v in 10; do echo -n "notebook variant $v: "; XINV_FUSED3_VARIANT=$v timeout 120 python scripts/prof_c3.py 50 300 300 602 | tail -1; done
