#!/bin/bash
for v in 0 2; do XINV_FUSED_GEN_VARIANT=$v python scripts/prof_c4.py > /tmp/o.txt 2>&1; echo "gen variant $v: $(head -1 /tmp/o.txt)"; done
for v in 3 9 7 2; do XINV_FUSED_RC_VARIANT=$v python scripts/prof_c4.py > /tmp/o.txt 2>&1; echo "rc variant $v: $(tail -1 /tmp/o.txt)"; done
